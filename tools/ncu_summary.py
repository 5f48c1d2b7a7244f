"""Developer helper: the metrics profiles/*_ncu_summaries.json keeps, out of an .ncu-rep
(`ncu -i rep --page raw --csv`).  Usage: python tools/ncu_summary.py rep [rep ...] > summaries.json
(keys = report stem; the metric list is the one the round-2 summaries use, plus warp execution efficiency)."""
import csv, io, json, subprocess, sys
from pathlib import Path

KEEP = """dram__bytes_read.sum dram__bytes_write.sum gpu__time_duration.sum
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum l1tex__t_sector_hit_rate.pct
l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum
launch__block_size launch__grid_size launch__registers_per_thread launch__shared_mem_per_block_dynamic
launch__shared_mem_per_block_static launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
lts__t_sector_hit_rate.pct lts__t_sectors.sum lts__t_bytes.sum lts__throughput.avg.pct_of_peak_sustained_elapsed
dram__throughput.avg.pct_of_peak_sustained_elapsed
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__inst_executed.sum smsp__thread_inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
smsp__issue_active.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__cycles_active.avg""".split()

out = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")]}
    for k in KEEP:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (vals[i] + " " + units[i]).strip()
    out[Path(rep).stem] = d
print(json.dumps(out, indent=1))

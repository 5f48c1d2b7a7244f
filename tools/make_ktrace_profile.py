"""Developer helper: profiles/k_trace_ncu_current.json (what bench.py reads for the ncu-only roofline fields) out of
an ncu --set full capture made by tools/prof_trace.sh and the hash of the library that was captured.
usage: python tools/make_ktrace_profile.py gpurun_out/<name>.ncu-rep gpurun_out/<name>.sha"""
import json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
rep, sha = sys.argv[1], Path(sys.argv[2]).read_text().split()[0]
summ = json.loads(subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), rep], capture_output=True, text=True).stdout)
t = summ[Path(rep).stem]
def byts(x):
    v, u = x.split()[:2]
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
out = {"capture": f"{Path(rep).name}: ncu --set full --clock-control none, tools/prof_trace.sh (tools/dev_prof.py NSENS=2048, second k_trace "
                  "launch = the 8.84 M-ray first-bounce wave of the 100 k-polygon office)",
       "rays_in_launch": 8842983,
       "dram_bytes_per_launch": byts(t["dram__bytes_read.sum"]) + byts(t["dram__bytes_write.sum"]),
       "thread_inst_per_inst": float(t["smsp__thread_inst_executed_per_inst_executed.ratio"]),
       "l2_hit_rate": float(t["lts__t_sector_hit_rate.pct"].split()[0]) / 100,
       "duration_ms_under_ncu": float(t["gpu__time_duration.sum"].split()[0]),
       "librb200_sha16": sha}
(ROOT / "profiles" / "k_trace_ncu_current.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out))

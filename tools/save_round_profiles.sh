#!/bin/bash
# developer helper: copy the records of tools/gpu_round_check.sh <tag> from gpurun_out/ into profiles/<round>_*
# usage: bash tools/save_round_profiles.sh <tag> <round>      e.g. t8 r6
t=$1; r=$2
cp gpurun_out/${t}_pytest.log profiles/${r}_pytest_gpu.log
cp gpurun_out/${t}_bench.json profiles/${r}_bench_line.json
cp gpurun_out/${t}_c1.json profiles/${r}_c1_line.json
cp gpurun_out/${t}_launches.csv profiles/${r}_launches_bench.csv
cp gpurun_out/${t}_smoke.log profiles/${r}_smoke.log
python tools/make_ktrace_profile.py gpurun_out/${t}_trace.ncu-rep gpurun_out/${t}_trace.sha > /dev/null
python tools/ncu_summary.py gpurun_out/${t}_trace.ncu-rep gpurun_out/${t}_shade_fast.ncu-rep > profiles/${r}_ncu_summaries.json
python - "$r" <<'PY'
import csv, collections, sys
r = sys.argv[1]
rows = [x for x in csv.reader(open(f'profiles/{r}_launches_bench.csv')) if len(x) > 5]
hdr = [x for x in rows if 'Kernel Name' in x][0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
dd = collections.defaultdict(lambda: [0, 0.0])
for x in rows[rows.index(hdr) + 1:]:
    try: v = float(x[vi].replace(',', ''))
    except ValueError: continue
    u = x[ui]; v = v / 1e3 if u == 'us' else v / 1e6 if u == 'ns' else v * 1e3 if u == 's' else v
    n = x[ki].split('(')[0]; dd[n][0] += 1; dd[n][1] += v
tot = sum(v[1] for v in dd.values())
with open(f'profiles/{r}_launch_shares.csv', 'w') as f:
    f.write("kernel,launches,total_ms,share\n")
    for k, v in sorted(dd.items(), key=lambda kv: -kv[1][1]): f.write(f"{k},{v[0]},{v[1]:.3f},{v[1] / tot:.4f}\n")
PY
sha256sum pyradiance_b200/librb200.so | cut -c1-16; grep -o '"librb200_sha16": "[0-9a-f]*"' profiles/${r}_bench_line.json; grep -o '"librb200_sha16": "[0-9a-f]*"' profiles/k_trace_ncu_current.json

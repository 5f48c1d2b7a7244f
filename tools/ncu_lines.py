"""Developer helper: per-source-line summary of an `ncu --page source --csv
--print-source cuda,sass` export (samples, instructions, live lanes, top stalls).
Usage: ncu -i rep --page source --csv --print-source cuda,sass > x.csv; python tools/ncu_lines.py x.csv [min_pct]"""
import csv, sys
from collections import defaultdict
csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
hdr = None; cur_file = None; cur_line = None; cur_src = ""
agg = defaultdict(lambda: defaultdict(float)); srcs = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr[2:], r[2:]))   # the sass-level columns (second "Source" shadows the first)
    if r[0] != "":
        cur_line = (cur_file, int(r[0])); srcs[cur_line] = r[1].strip()
    if cur_line is None or not r[2]: continue
    a = agg[cur_line]
    def f(k):
        try: return float(d.get(k, 0) or 0)
        except ValueError: return 0.0
    a["samples"] += f("# Samples"); a["inst"] += f("Instructions Executed"); a["tinst"] += f("Thread Instructions Executed")
    for k in d:
        if k.startswith("stall_") and "(Not Issued)" not in k: a[k] += f(k)
tot_s = sum(a["samples"] for a in agg.values()); tot_i = sum(a["inst"] for a in agg.values()); tot_t = sum(a["tinst"] for a in agg.values())
print(f"total samples {tot_s:.0f} warp-inst {tot_i:.3e} thread-inst {tot_t:.3e} avg lanes {tot_t/max(tot_i,1):.1f}")
st = defaultdict(float)
for a in agg.values():
    for k, v in a.items():
        if k.startswith("stall_"): st[k] += v
print("stalls:", ", ".join(f"{k[6:]} {100*v/tot_s:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
print(f"{'file:line':24s} {'smp%':>5s} {'inst%':>5s} {'lanes':>5s}  top stalls / source")
for key, a in sorted(agg.items(), key=lambda x: -x[1]["samples"]):
    p = 100 * a["samples"] / tot_s
    if p < minpct: break
    top = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_")), key=lambda x: -x[1])[:3]
    print(f"{key[0]+':'+str(key[1]):24s} {p:5.1f} {100*a['inst']/tot_i:5.1f} {a['tinst']/max(a['inst'],1):5.1f}  " +
          " ".join(f"{k}={100*v/max(a['samples'],1):.0f}%" for k, v in top) + "  | " + srcs.get(key, "")[:90])

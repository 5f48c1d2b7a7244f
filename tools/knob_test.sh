#!/bin/bash
# developer helper (GPU box): k_trace occupancy / carve-out knobs on the bench scene
export RB_TMP=/tmp/rbt
NSENS=512 python tools/dev_prof.py > /dev/null 2>&1
for ev in " " "RB_TRACE_CTAS_PER_SM=4" "RB_TRACE_CTAS_PER_SM=4 RB_TRACE_CARVEOUT=55" "RB_TRACE_CARVEOUT=70" "RB_TRACE_CARVEOUT=75" "RB_TRACE_CARVEOUT=90" "RB_TRACE_CTAS_PER_SM=3 RB_TRACE_CARVEOUT=45"; do
  echo "== $ev"; env $ev RB_DEBUG_GRID=1 NSENS=4096 REPS=2 python tools/dev_prof.py 2>&1 | grep -E "k_trace:|rep 1" | sort -u | cut -c1-170
done

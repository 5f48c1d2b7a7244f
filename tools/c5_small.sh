#!/bin/bash
# developer helper (GPU box): configs[4] with 20000 sensors on one GPU, for every library given
for so in "$@"; do
  echo "== $so"
  RB200_LIBRARY=$PWD/$so python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29532 tools/target_c5.py --sensors 20000 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('device_ms_max_over_ranks','rays_total','rays_per_sec','k_trace_share')}, d['checks']['mean_row_sum'], d['checks'].get('reference_rows'))"
done

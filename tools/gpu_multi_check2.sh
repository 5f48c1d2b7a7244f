#!/bin/bash
# developer helper (8-GPU box): 8-GPU bench line and configs[4] at full size.  usage: bash tools/gpu_multi_check2.sh <tag>
t=${1:-t13}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 \
    > gpurun_out/${t}_bench_8gpu.json 2> gpurun_out/${t}_bench_8gpu.err; cut -c1-600 gpurun_out/${t}_bench_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tools/target_c5.py \
    --out gpurun_out/${t}_target_c5.json 2>&1 | tail -1 | cut -c1-900

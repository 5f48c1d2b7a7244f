#!/bin/bash
# developer helper (8-GPU box): the multi-GPU records of a round under gpurun_out/<tag>_*: GPU tests (the 2-GPU ones
# included), the 8-GPU bench line and the north-star target run.  usage: bash tools/gpu_multi_check.sh <tag>
t=${1:-t5}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${t}_topo.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${t}_pytest_8gpu.log 2>&1; tail -3 gpurun_out/${t}_pytest_8gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 \
    > gpurun_out/${t}_bench_8gpu.json 2> gpurun_out/${t}_bench_8gpu.err; cat gpurun_out/${t}_bench_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/target_c3.py \
    --out gpurun_out/${t}_target_c3_8gpu.json > gpurun_out/${t}_target_c3.log 2>&1; tail -2 gpurun_out/${t}_target_c3.log

#!/bin/bash
# developer helper (GPU box): everything a round's records need from one fresh box, under gpurun_out/<tag>_*:
# GPU tests, smoke, the bench line, the configs[0] line, the ncu launch list of bench.py and full captures of k_trace
# and k_shade_fast.  usage: bash tools/gpu_round_check.sh <tag>
t=${1:-t1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${t}_pytest.log 2>&1; tail -3 gpurun_out/${t}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${t}_smoke.log 2>&1; tail -1 gpurun_out/${t}_smoke.log
python bench.py > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; cat gpurun_out/${t}_bench.json
python bench.py --config c1 > gpurun_out/${t}_c1.json 2> gpurun_out/${t}_c1.err; cat gpurun_out/${t}_c1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${t}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${t}_ncu_bench.log 2>&1
bash tools/prof_trace.sh ${t}_trace
bash tools/prof_kernel.sh ${t}_shade_fast k_shade_fast 1

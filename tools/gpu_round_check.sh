set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t1_pytest.log 2>&1; tail -3 gpurun_out/t1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t1_smoke.log 2>&1; tail -1 gpurun_out/t1_smoke.log
python bench.py > gpurun_out/t1_bench.json 2> gpurun_out/t1_bench.err; cat gpurun_out/t1_bench.json
python bench.py --config c1 > gpurun_out/t1_c1.json 2> gpurun_out/t1_c1.err; cat gpurun_out/t1_c1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/t1_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/t1_ncu_bench.log 2>&1
bash tools/prof_trace.sh t1_trace
bash tools/prof_kernel.sh t1_shade_fast k_shade_fast 1

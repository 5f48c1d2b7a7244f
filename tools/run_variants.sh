#!/bin/bash
# developer helper: time tools/dev_prof.py against every variants_*.so in the repo root
export RB_TMP=gpurun_out/tmp
for so in pyradiance_b200/librb200.so variants_*.so; do
  echo "== $so"
  RB200_LIBRARY=$PWD/$so NSENS=${NSENS:-4096} REPS=2 python tools/dev_prof.py 2>&1 | tail -1
done
rm -rf gpurun_out/tmp

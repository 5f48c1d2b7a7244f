#!/bin/bash
# developer helper: time tools/dev_prof.py against every variants_*.so in the repo root
# (VARIANT_ENVS="A=1 B=2;A=3" repeats each library under those environments)
export RB_TMP=gpurun_out/tmp
IFS=';' read -ra ENVS <<< "${VARIANT_ENVS:- }"
for so in pyradiance_b200/librb200.so variants_*.so; do
  for ev in "${ENVS[@]}"; do
    echo "== $so $ev"
    env $ev RB200_LIBRARY=$PWD/$so NSENS=${NSENS:-4096} REPS=2 python tools/dev_prof.py 2>&1 | tail -1
  done
done
rm -rf gpurun_out/tmp

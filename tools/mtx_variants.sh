#!/bin/bash
# developer helper (GPU box): tools/dev_mtx.py against every variants_*.so (NI=145 and NI=2305)
for so in pyradiance_b200/librb200.so variants_*.so; do
  echo "== $so"
  RB200_LIBRARY=$PWD/$so timeout 200 python tools/dev_mtx.py 2>&1 | grep -E "run 3|max rel"
  RB200_LIBRARY=$PWD/$so NR=20000 NI=2305 timeout 200 python tools/dev_mtx.py 2>&1 | grep -E "run 3|max rel"
done

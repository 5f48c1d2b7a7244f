#!/bin/bash
# developer helper (GPU box): ncu --set full capture of the big k_trace wave of tools/dev_prof.py -> gpurun_out/<name>.ncu-rep
# usage: tools/prof_trace.sh <name> [library.so]
name=${1:-trace}; lib=${2:-pyradiance_b200/librb200.so}
export RB_TMP=/tmp/rbt
sha256sum $lib | cut -c1-16 > gpurun_out/$name.sha
NSENS=2048 python tools/dev_prof.py > /dev/null 2>&1    # builds the scene once
RB200_LIBRARY=$PWD/$lib NSENS=2048 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 \
   -o gpurun_out/$name -f python tools/dev_prof.py 2>&1 | tail -3

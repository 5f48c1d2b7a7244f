#!/bin/bash
# developer helper (GPU box): ncu --set full capture of one launch of a kernel of tools/dev_prof.py
# usage: tools/prof_kernel.sh <name> <kernel regex> <skip> [library.so]
name=$1; kern=$2; skip=${3:-1}; lib=${4:-pyradiance_b200/librb200.so}
export RB_TMP=/tmp/rbt
NSENS=2048 python tools/dev_prof.py > /dev/null 2>&1
RB200_LIBRARY=$PWD/$lib NSENS=2048 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 \
   -o gpurun_out/$name -f python tools/dev_prof.py 2>&1 | tail -2

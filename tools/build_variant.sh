#!/bin/bash
# developer helper: tools/build_variant.sh <name> [-DMACRO=val ...] -> variants_<name>.so (engine rebuilt with the macros)
set -e
name=$1; shift
cd "$(dirname "$0")/../pyradiance_b200/csrc"
make -s all >/dev/null
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -fPIC,-O2 -Xptxas -v --expt-relaxed-constexpr "$@" -c rb_engine.cu -o /tmp/rb_engine_$name.o 2> /tmp/rb_engine_$name.log
grep -A2 "k_trace" /tmp/rb_engine_$name.log | grep -E "spill|registers" | tr '\n' ' '; echo
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants_$name.so /tmp/rb_engine_$name.o rb_api.o rb_scene.o rb_octbuild.o rb_format.o rb_mtx.o rb_mtx_tc.o rb_octbuild_gpu.o rb_views.o rb_bsdf.o

#!/bin/bash
# Debug build of the library with -DPICARD_RB_TRACE (phase timestamps of every warp of CTA 0 in rb_loss_kernel<128>);
# run here (no GPU needed):  bash profiles/rb_trace.sh build      -> build/trace/libpicard_b200.so
# on the GPU box:            bash profiles/rb_trace.sh run [skew] -> gpurun_out/rb_trace_<skew>.json
set -e
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  mkdir -p build/trace
  make -C picard-ica_b200/csrc -j8 OBJDIR=../../build/obj_trace LIB=../../build/trace/libpicard_b200.so \
    NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -I/usr/include -DPICARD_RB_TRACE" >/dev/null
  ls -la build/trace/libpicard_b200.so
else
  skew=${2:-0}
  PICARD_RB_SKEW=$skew PICARD_B200_LIB=$PWD/build/trace/libpicard_b200.so python profiles/rb_trace.py gpurun_out/rb_trace_$skew.json
fi

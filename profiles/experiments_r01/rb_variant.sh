#!/bin/bash
# A/B builds of the library with one extra -D flag (kernel geometry experiments).
#   here:        bash profiles/rb_variant.sh build <name> <-DFLAG ...>   -> build/variant_<name>/libpicard_b200.so
#   on the box:  bash profiles/rb_variant.sh run <name> <pass_bench args>
set -e
cd "$(dirname "$0")/.."
cmd=$1; name=$2; shift 2
if [ "$cmd" = build ]; then
  mkdir -p build/variant_$name
  make -C picard-ica_b200/csrc -j8 OBJDIR=../../build/obj_$name LIB=../../build/variant_$name/libpicard_b200.so \
    NVFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -I/usr/include $*" >/dev/null
  ls -la build/variant_$name/libpicard_b200.so
else
  PICARD_B200_LIB=$PWD/build/variant_$name/libpicard_b200.so timeout -s KILL 90 python profiles/pass_bench.py "$@"
fi

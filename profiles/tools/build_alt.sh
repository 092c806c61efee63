#!/bin/bash
# bash profiles/tools/build_alt.sh NAME [-DFLAG=..]...   -- builds rust-sloth_b200/alt/lib_NAME.so with extra defines
# (same flags as the Makefile otherwise); the A/B scripts select a build through SLOTH_B200_LIB.
set -e
cd "$(dirname "$0")/../../rust-sloth_b200"
name=$1; shift
mkdir -p alt
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall -Xptxas -v "$@" -shared -o alt/lib_$name.so csrc/capi.cu host/mesh_io.cpp host/wire.cpp -lpthread 2>&1 \
  | grep -A2 "k_triILb0ELb0ELb1E" | grep -E "registers|spill" || true
ls -la alt/lib_$name.so

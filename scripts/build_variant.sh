#!/bin/bash
# scripts/build_variant.sh <name> <tu> [extra nvcc flags...]  →  build/libagx_<name>.so
# Tuning variants for scripts/kbench.py (lib=...) and scripts/enc_bench.py (AGX_LIB=...): recompiles ONE translation unit
# (e.g. agx_cnn, agx_step_hovering) with the extra flags and links it against the other objects of the regular build
# (run `python __graft_entry__.py` first so that build/obj/ is current).
set -e
cd "$(dirname "$0")/.."
name=$1; tu=$2; shift 2
mkdir -p build/obj_variant
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -prec-div=false -prec-sqrt=false \
  -Iinclude -Iairgym_b200/csrc -Xcompiler -fPIC "$@" -c -o build/obj_variant/${tu}_$name.o airgym_b200/csrc/$tu.cu
others=$(ls build/obj/*.o | grep -v "/$tu.o")
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/libagx_$name.so build/obj_variant/${tu}_$name.o $others
echo build/libagx_$name.so

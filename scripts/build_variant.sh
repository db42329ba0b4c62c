#!/bin/bash
# scripts/build_variant.sh <name> [extra nvcc flags...]  →  build/libagx_<name>.so (tuning variants for scripts/kbench.py lib=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -prec-div=false -prec-sqrt=false \
  -Iinclude -Iairgym_b200/csrc -shared -Xcompiler -fPIC "$@" -o build/libagx_$name.so airgym_b200/csrc/agx_step.cu airgym_b200/csrc/agx_ppo.cu airgym_b200/csrc/agx_mlp.cu

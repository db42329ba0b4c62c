#!/bin/bash
# scripts/ab_encoder_branches.sh [branch...]  →  build/libagx_<branch>.so for each encoder branch (default: the two next/ branches)
# Builds the depth-encoder translation unit of a git branch (sources read straight from the branch with `git show`, nothing is
# checked out) against the regular build's other objects, so that one GPU call can A/B them:
#   python scripts/enc_bench.py --n 8192 --skip_cudnn --libs build/libagx_next_encoder-v3.so,build/libagx_next_encoder-tc.so
# (enc_bench reports max abs error against torch fp32 next to the time, so the run doubles as the variants' first GPU parity check).
set -e
cd "$(dirname "$0")/.."
branches=("$@"); [ ${#branches[@]} -eq 0 ] && branches=(next/encoder-v3 next/encoder-tc)
python __graft_entry__.py > /dev/null   # build/obj/ current
for br in "${branches[@]}"; do
  name=$(echo "$br" | tr '/' '_')
  src=build/src_$name; mkdir -p "$src" build/obj_variant
  for f in agx_cnn.cu agx_cnn.cuh; do git show "$br:airgym_b200/csrc/$f" > "$src/$f"; done
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -prec-div=false -prec-sqrt=false \
    -Iinclude -I"$src" -Iairgym_b200/csrc -Xcompiler -fPIC -c -o build/obj_variant/agx_cnn_$name.o "$src/agx_cnn.cu"
  others=$(ls build/obj/*.o | grep -v "/agx_cnn.o")
  /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/libagx_$name.so build/obj_variant/agx_cnn_$name.o $others
  echo build/libagx_$name.so
done

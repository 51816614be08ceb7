#!/bin/bash
# tools/build_variant.sh NAME [-DMODK_...=..]...  -> gpurun_out/variants/libmod_NAME.so  (kernel tuning builds)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
mkdir -p "$ROOT/variants"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart static -shared \
  -I "$ROOT/include" -I "$ROOT/modulate_b200/csrc" "$@" -o "$ROOT/variants/libmod_$NAME.so" \
  "$ROOT"/modulate_b200/csrc/*.cu $(ls "$ROOT"/modulate_b200/csrc/*.cpp 2>/dev/null)
echo "$ROOT/variants/libmod_$NAME.so"

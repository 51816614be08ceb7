#!/bin/bash
# Build tuning variants of the library into variants/ (git-ignored; they travel to the GPU box).
#   tools/build_variants.sh name "-DMODK_THREADS=256 -DMODK_UNROLL=2" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart static -shared \
        $flags -I include -I modulate_b200/csrc -o variants/libmod_$name.so \
        modulate_b200/csrc/*.cpp modulate_b200/csrc/*.cu &
done
wait
ls -la variants

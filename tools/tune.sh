#!/bin/bash
# tools/tune.sh -- time every kernel variant under variants/ with the kernel-only bench (on the GPU box)
cd "$(dirname "$0")/.."
for lib in variants/libmod_*.so; do
  echo -n "$(basename $lib) cfg2: "
  MODULATE_B200_LIB=$PWD/$lib python bench.py --steps 30 --warmup 5 --kernel-only 2>&1 | tail -n 1
done

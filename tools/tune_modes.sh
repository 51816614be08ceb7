#!/bin/bash
cd "$(dirname "$0")/.."
for mode in 0 1; do for lib in variants/libmod_*.so; do
  echo -n "mode=$mode $(basename $lib): "
  MOD_GRID_MODE=$mode MODULATE_B200_LIB=$PWD/$lib python bench.py --steps 30 --warmup 5 --kernel-only 2>&1 | tail -n 1 | cut -c1-110
done; done

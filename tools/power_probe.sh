#!/bin/bash
# tools/power_probe.sh -- power / clocks while a sustained kernel runs (~3 s each)
cd "$(dirname "$0")/.."
probe() {
  nvidia-smi --query-gpu=power.draw,clocks.sm,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 100 > /tmp/pw.csv &
  local pid=$!
  "$@"
  kill $pid; wait $pid 2>/dev/null
  awk -F, '{p=$1+0; c=$2+0; if(p>mp)mp=p; sum+=c; n++; if($3 ~ /Active/ && $3 !~ /Not/) cap++} END{printf "   power max %.0f W, mean SM clock %.0f MHz, sw_power_cap in %d of %d samples\n", mp, sum/n, cap, n}' /tmp/pw.csv
}
probe tools/copy_skeleton_bench M 9000
sleep 5
probe tools/copy_skeleton_bench B 9000
sleep 5
probe tools/copy_skeleton_bench C 9000
sleep 5
probe python bench.py --steps 8000 --warmup 5 --kernel-only

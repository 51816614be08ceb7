#!/bin/bash
# A/B of kernel variants on the GPU box: tools/gpu_ab.sh "<variant names>" "<workloads>"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/ab_$(date +%H%M%S).log
for v in $1; do
    for w in $2; do
        lib=variants/libmod_$v.so
        [ "$v" = "default" ] && lib=modulate_b200/libmodulate_b200.so
        echo -n "== $v $w  " >> $out
        MODULATE_B200_LIB=$PWD/$lib timeout 300 python bench.py --workload $w --kernel-only --steps 20 --warmup 3 2>&1 | tail -1 | \
          python -c "import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('payload_gbs %.1f frac %.3f kernel_ms %.4f value %.1f parity %d' % (r['payload_gbs'], r['frac'], r['kernel_ms'], d['value'], d['parity_bytes_checked']))
except Exception as e:
    print('FAILED', e)" >> $out
    done
done
cat $out

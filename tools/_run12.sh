for pass in 1 2; do for v in coal10 default; do
lib=variants/libmod_$v.so; [ "$v" = "default" ] && lib=modulate_b200/libmodulate_b200.so
echo -n "$v: "
MODULATE_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-extras --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; s=r['sustained']
print('burst %.1f GB/s frac %.3f | sustained %.1f GB/s frac %.3f at %.0f MHz %.0f W (mean %.0f W) | e2e %.1f' % (r['payload_gbs'], r['frac'], s['payload_gbs'], s['frac'], s['clocks']['sm_mhz'], s['clocks']['power_w_max'], s['clocks']['power_w_mean'], d['e2e']['value']))"
done; done

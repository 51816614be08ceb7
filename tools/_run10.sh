timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest4.log
for rc in 0 512 128; do for w in cfg4 cfg2; do
echo -n "MOD_REC_CHUNKS=$rc $w  "
MOD_REC_CHUNKS=$rc timeout 300 python bench.py --workload $w --kernel-only --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('payload_gbs %.1f frac %.3f kernel_ms %.4f parity %d' % (r['payload_gbs'], r['frac'], r['kernel_ms'], d['parity_bytes_checked']))"
done; done

timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
t0=$(date +%s); python bench.py > gpurun_out/bench_r2e.log 2>&1; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
grep "^{" gpurun_out/bench_r2e.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value %.0f'%d['value'],'ms/step %.4f'%d['ms_per_step'],'kernel_ms %.4f'%d['roofline']['kernel_ms'],'frac %.3f'%d['roofline']['frac'],'sustained %.3f'%d['roofline']['sustained']['frac'],'e2e %.1f'%d['e2e']['value'])
for w in ('cfg2','cfg4'):
    x=d['extra'][w]; print(w,'steps',x['steps'],'frac %.3f'%x['roofline']['frac'],'sustained %.3f'%x['roofline']['sustained']['frac'],x['roofline']['sustained']['clocks']['sm_mhz'],x['clocks']['sm_mhz'],x['clocks']['reasons'])
x=d['extra']['cfg5']; print('cfg5 unpack %.1f pack %.1f GB/s'%(x['unpack_gbs'],x['pack_gbs']), x['samples_s'])"
python bench.py --kernel-only --workload cfg2 --steps 3 --warmup 3 > gpurun_out/plain_k8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cycle_batch_kernel -s 4 -c 2 -o gpurun_out/r02_cfg2_final2 \
    python bench.py --kernel-only --workload cfg2 --steps 3 --warmup 3 > gpurun_out/ncu_f8.log 2>&1; echo "ncu full cfg2 rc=$?"

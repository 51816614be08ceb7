nproc
for i in 1 2; do
MOD_TRACE=1 timeout 600 python bench.py --workload cfg5 > gpurun_out/bench_cfg5_c.log 2>&1; echo "cfg5 rc=$?"; grep "^\[mod\] Extract\|^\[mod\] Save\|PARITY" gpurun_out/bench_cfg5_c.log | grep -v "setup:\|cleanup:" | tail -3; grep "^{" gpurun_out/bench_cfg5_c.log | python -c "import sys,json; d=json.loads(sys.stdin.read())['e2e']; print({k:round(d[k],3) for k in ('value','unpack_s','pack_s','unpack_gbs','pack_gbs')})"
done
timeout 900 python -m pytest tests/test_gpu_facade.py -x -q 2>&1 | tail -3

nproc; free -g | head -2; df -h /dev/shm | tail -1
timeout 900 python -m pytest tests/test_gpu_facade.py tests/test_gpu_multidev.py -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest3.log
MOD_TRACE=1 timeout 600 python bench.py --workload cfg5 > gpurun_out/bench_cfg5.log 2>&1; echo "cfg5 rc=$?"; grep "^\[mod\]\|^{" gpurun_out/bench_cfg5.log | tail -8 | cut -c1-900

nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_facade.py -x -q > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; grep "^{" gpurun_out/bench_n2.log | tail -1 | cut -c1-1500
timeout 300 python tools/pcie_probe.py --gpus 2 --json gpurun_out/pcie_2.json; echo "probe rc=$?"
MOD_TRACE=1 timeout 600 python bench.py --workload cfg5 > gpurun_out/bench_cfg5_2gpu.log 2>&1; echo "cfg5 rc=$?"; grep "^\[mod\]\|^{\|PARITY" gpurun_out/bench_cfg5_2gpu.log | tail -8 | cut -c1-1200

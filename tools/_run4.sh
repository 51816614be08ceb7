timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2b.log 2>&1; echo "bench rc=$?"
python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/plain_k3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cycle_|build_tiles" -c 60 --csv --log-file gpurun_out/r02_launches_cfg3.csv \
    python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_l.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:cycle_batch_kernel -s 4 -c 2 -o gpurun_out/r02_cfg3_full \
    python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_f3.log 2>&1; echo "ncu full cfg3 rc=$?"
python bench.py --kernel-only --workload cfg2 --steps 3 --warmup 3 > gpurun_out/plain_k2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cycle_batch_kernel -s 4 -c 2 -o gpurun_out/r02_cfg2_full \
    python bench.py --kernel-only --workload cfg2 --steps 3 --warmup 3 > gpurun_out/ncu_f2.log 2>&1; echo "ncu full cfg2 rc=$?"
ls -la gpurun_out/*.ncu-rep

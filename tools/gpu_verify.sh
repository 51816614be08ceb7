#!/bin/bash
# tools/gpu_verify.sh -- one gpurun call that re-verifies a build on a B200: GPU tests, smoke(), the driver-style bench
# (both arms), the ncu launch list and one ncu --set full capture per workload (each after its own plain run exited 0).
#   gpurun --timeout 1500 -- bash tools/gpu_verify.sh

timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
t0=$(date +%s); python bench.py > gpurun_out/bench_final_n1.log 2>&1; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
t0=$(date +%s); python bench.py --impl reference > gpurun_out/bench_ref_final.log 2>&1; echo "ref rc=$? wall=$(( $(date +%s) - t0 ))s"
python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/plain_k5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cycle_|build_tiles" -c 60 --csv --log-file gpurun_out/r02_launches_cfg3_final.csv \
    python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_l5.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:cycle_batch_kernel -s 4 -c 2 -o gpurun_out/r02_cfg3_final \
    python bench.py --kernel-only --steps 3 --warmup 3 > gpurun_out/ncu_f5.log 2>&1; echo "ncu full cfg3 rc=$?"
python bench.py --kernel-only --workload cfg2 --steps 3 --warmup 3 > gpurun_out/plain_k6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cycle_batch_kernel -s 4 -c 2 -o gpurun_out/r02_cfg2_final \
    python bench.py --kernel-only --workload cfg2 --steps 3 --warmup 3 > gpurun_out/ncu_f6.log 2>&1; echo "ncu full cfg2 rc=$?"
python bench.py --kernel-only --workload cfg4 --steps 3 --warmup 3 > gpurun_out/plain_k7.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cycle_batch_kernel -s 4 -c 1 -o gpurun_out/r02_cfg4_final \
    python bench.py --kernel-only --workload cfg4 --steps 3 --warmup 3 > gpurun_out/ncu_f7.log 2>&1; echo "ncu full cfg4 rc=$?"
ls -la gpurun_out/*final*.ncu-rep
grep "^{" gpurun_out/bench_final_n1.log | tail -1 | cut -c1-600

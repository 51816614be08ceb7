nvidia-smi -L > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
tools/gpu_ab.sh "r1 t128u4c6 t128u4c8 t128u4c5 t128u4c4 t256u2 t256u4 t64u4 copyonly" "cfg2 cfg3 cfg4"

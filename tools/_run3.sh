tools/gpu_ab.sh "p2k c10 s1c8 s1c10 s1c12 s2c8 s2c10 s2c12 s2c16 s2t64c20 s1t64c20" "cfg2 cfg3"
tools/gpu_ab.sh "c10 s2c10 s1c10" "cfg4"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_r2a.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r2a.log 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref_r2a.log

#!/bin/bash
# tools/gpu_scaling.sh -- strong scaling of the headline workload on one multi-GPU box, the way the driver launches it:
#   gpurun --gpus 8 --timeout 900 -- bash tools/gpu_scaling.sh "2 4 8"
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
for n in ${1:-2 4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_final_n$n.log 2>&1; echo "bench n$n rc=$?"
  grep "^{" gpurun_out/bench_final_n$n.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value %.0f'%d['value'],'ms/step %.3f'%d['ms_per_step'],'frac %.3f'%d['roofline']['frac'],'sustained %.3f'%d['roofline']['sustained']['frac'],'parity',d['parity_bytes_checked'],'e2e %.1f'%d['e2e']['value'],'inproc',d['extra'].get('e2e_inprocess'))"
done

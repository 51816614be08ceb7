timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multidev.py -x -q 2>&1 | tail -3
bash tools/_run12.sh
tools/gpu_ab.sh "coal10 default coal10 default" "cfg4 cfg2"

timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/cli_timing.sh 2>&1 | tail -8

set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "async or h2 or batch" 2>&1 | tail -3
python tools/e2e_scaling.py 2>&1 | grep zero_copy

set -x
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -q -x -k "grouped or collection or plan_options or evaluate_variants or async" --durations=5 2>&1 | tail -12

set -x
timeout 900 python -m pytest tests/test_sampler.py tests/test_gpu_dropin.py -q -x --durations=6 2>&1 | tail -14

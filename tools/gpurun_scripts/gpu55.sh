timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_sampler.py -q -m gpu 2>&1 | tail -3

set -x
timeout 80 python -m pytest tests/test_sampler.py -q -m gpu -x 2>&1 | tail -4

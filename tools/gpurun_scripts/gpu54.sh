timeout 300 python -m pytest tests/test_sampler.py -q -m gpu 2>&1 | tail -3
python tools/sanitize_small.py 2>&1 | tail -2

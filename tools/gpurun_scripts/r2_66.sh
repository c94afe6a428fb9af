timeout 42 python -m pytest tests/test_facade_surface.py -q -m gpu -x 2>&1 | tail -12

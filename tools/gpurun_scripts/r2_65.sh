timeout 50 python -m pytest tests/test_trajectories.py -q -m gpu -k "wide_mid" 2>&1 | tail -4

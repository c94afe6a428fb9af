set -x
timeout 600 python -m pytest tests/test_trajectories.py -q -m gpu -x --durations=5 2>&1 | tail -25

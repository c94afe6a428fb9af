set -x
timeout 1800 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -25

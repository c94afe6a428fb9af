set -x
timeout 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -25
echo done

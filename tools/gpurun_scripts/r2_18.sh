set -x
DMX_WIDE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "streaming or hea10 or lih12 or eleven or ten_qubit or twelve" 2>&1 | tail -4

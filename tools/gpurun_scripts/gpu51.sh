python tools/host_call_scaling.py 2>&1 | tail -7

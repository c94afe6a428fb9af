python tools/qem_breakdown.py 2>&1 | tail -6

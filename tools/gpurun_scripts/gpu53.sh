timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py 2>&1 | tail -8

set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x -k "h2 or golden or batch or async or grouped" 2>&1 | tail -3
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps, extra=()):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "5", "--no-cpu-baseline", "--no-also", *extra] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        e2e = d.get("e2e") or {}
        print(f"{wl:9s} {str(opts):20s} value {d['value']:.5g}  ms {d['ms_per_step']:.5f} parity {d['parity_max_abs_err']} e2e {e2e.get('value',0):.4g} sync {(e2e.get('synchronous') or {}).get('value',0):.4g}", flush=True)
    except Exception as e:
        print(wl, opts, "FAILED", r.stdout[-300:], r.stderr[-1500:], flush=True)
run("h2_sweep", [], 200)
run("h2_sweep", [], 200)
run("h2_bonds", [], 100)
PY
python tools/e2e_scaling.py 2>&1 | grep zero_copy

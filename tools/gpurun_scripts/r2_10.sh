set -x
export DMX_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "streaming or hea10 or lih12_26 or eleven" 2>&1 | tail -3
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps, extra=()):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "3", "--no-cpu-baseline", "--no-also", *extra] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        e2e = d.get("e2e") or {}
        print(f"{wl:9s} {str(opts):20s} value {d['value']:.5g}  ms {d['ms_per_step']:.4f} parity {d['parity_max_abs_err']} e2e {e2e.get('value',0):.4g} frac {d['roofline']['frac']:.3f} est {d.get('estimator',{}).get('value')} dedup {d.get('dedup',{}).get('value')}", flush=True)
        if wl == "swap7_qem": print(json.dumps(d["e2e"])[:900]); print(json.dumps(d.get("estimator")))
    except Exception as e:
        print(wl, opts, "FAILED", r.stdout[-300:], r.stderr[-1500:], flush=True)
run("hea10", [], 4, ("--no-e2e",))
run("lih12", [], 2, ("--no-e2e",))
run("swap7_qem", [], 5)
PY

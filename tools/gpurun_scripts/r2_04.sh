set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x -k "not streaming and not eleven and not lih12 and not hea10 and not twelve and not ten_qubit and not similarity" 2>&1 | tail -5
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps, extra=()):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "5", "--no-cpu-baseline", "--no-also", *extra] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        e2e = d.get("e2e") or {}
        print(f"{wl:8s} {str(opts):28s} value {d['value']:.5g}  ms {d['ms_per_step']:.5f} parity {d['parity_max_abs_err']} e2e {e2e.get('value',0):.4g} sync {(e2e.get('synchronous') or {}).get('value',0):.4g} frac {d['roofline']['frac']:.3f}", flush=True)
    except Exception as e:
        print(wl, opts, "FAILED", r.stderr[-800:], flush=True)
run("h2_sweep", [], 100)
run("h2_sweep", ["frame32s=0"], 100)
run("h2_sweep", [], 100)
run("hea10", ["scale_smem=1"], 4, ("--no-e2e",))
run("hea10", [], 4, ("--no-e2e",))
PY
python tools/batch_scaling.py 2>&1 | tail -9

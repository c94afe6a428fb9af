set -x
export DMX_VERBOSE=1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "streaming or hea10 or lih12_26 or eleven" 2>&1 | tail -3
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps, extra=()):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "3", "--no-cpu-baseline", "--no-also", "--no-e2e", *extra] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{wl:9s} {str(opts):20s} value {d['value']:.5g}  ms {d['ms_per_step']:.4f} parity {d['parity_max_abs_err']} fexec {d['roofline']['frac_executed']:.3f}", flush=True)
    except Exception as e:
        print(wl, opts, "FAILED", r.stdout[-300:], r.stderr[-1500:], flush=True)
run("hea10", [], 4)
run("lih12", [], 2)
run("hea10", ["wide=1"], 4)
PY

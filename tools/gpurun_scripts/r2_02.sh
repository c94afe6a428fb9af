set -x
export DMX_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "streaming or hea10 or lih12 or eleven or ten_qubit or twelve or plan_options" 2>&1 | tail -8
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "3", "--no-cpu-baseline", "--no-e2e"] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{wl:6s} {str(opts):40s} value {d['value']:.5g}  ms {d['ms_per_step']:.3f}  fexec {d['roofline']['frac_executed']:.3f} parity {d['parity_max_abs_err']}", flush=True)
    except Exception as e:
        print(wl, opts, "FAILED", r.stderr[-600:], flush=True)
    for l in r.stderr.splitlines():
        if l.startswith("[dmx]"): print("   ", l, flush=True); break
for opts in ([], ["bulk=2"], ["bulk=0"], ["direct_store=0"], ["bulk=0", "direct_store=0"], ["bulk=2", "direct_store=0"]):
    run("hea10", opts, 4)
for opts in ([], ["bulk=2"], ["bulk=0", "direct_store=0"]):
    run("lih12", opts, 2)
PY

set -x
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_sampler.py -q -x -k "variant or qp or qem or mitig or evaluate or sampl or streams" 2>&1 | tail -4
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps=20):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "5", "--no-cpu-baseline", "--no-also"] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        iso = d.get("isolated", {}).get("value", 0)
        est = d.get("estimator", {}).get("value", 0)
        print(f"{wl:9s} {str(opts):22s} value {d['value']:.4g} us {1e3*d['ms_per_step']:.2f} isolated {iso:.4g} e2e {d['e2e']['value']:.4g} dedup {d.get('dedup',{}).get('value',0):.4g} est {est:.4g} parity {d['parity_max_abs_err']}", flush=True)
    except Exception as e:
        print(wl, opts, "FAILED", r.stdout[-300:], r.stderr[-1500:], flush=True)
run("h2_qem", [])
run("h2_qem", [], 100)
PY

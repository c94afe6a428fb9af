set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x -k "h2 or host or async or batch or group or option or collection" 2>&1 | tail -3
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps=50):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "5", "--no-cpu-baseline", "--no-also"] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        ov = d.get("overlapped", {}).get("value", 0)
        print(f"{wl:9s} {str(opts):22s} value {d['value']:.4g} us {1e3*d['ms_per_step']:.2f} overlapped {ov:.4g} e2e {d['e2e']['value']:.4g} sync {d['e2e'].get('synchronous',{}).get('value',0):.4g} parity {d['parity_max_abs_err']}", flush=True)
    except Exception as e:
        print(wl, opts, "FAILED", r.stdout[-300:], r.stderr[-1500:], flush=True)
for o in ("1", "0", "40", "48", "56", "64", "101", "140", "148", "156", "164", "201", "248", "256", "264", "16"):
    run("h2_sweep", ["frame32s=" + o])
run("h2_bonds", [], 20)
run("h2_bonds", ["frame32s=101"], 20)
run("h2_bonds", ["frame32s=0"], 20)
PY

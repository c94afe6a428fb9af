set -x
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "3", "--no-cpu-baseline", "--no-also", "--no-e2e"] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{wl:9s} {str(opts):34s} value {d['value']:.5g}  ms {d['ms_per_step']:.3f} parity {d['parity_max_abs_err']} passes {d['config']['plan']['n_passes']}", flush=True)
    except Exception as e:
        print(wl, opts, "FAILED", r.stdout[-300:], r.stderr[-800:], flush=True)
run("hea10", ["tile_qubits=7", "triples=0"], 3)
run("hea10", ["tile_qubits=6", "triples=0"], 3)
run("lih12", ["triples=0"], 2)
PY

set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q -x -k "h2 or host or async or batch or group or option or collection" 2>&1 | tail -3
python - <<'PY'
import subprocess, json, sys
def run(wl, opts, steps=50):
    cmd = [sys.executable, "bench.py", "--workload", wl, "--steps", str(steps), "--warmup", "5", "--no-cpu-baseline", "--no-also"] + sum([["--opt", o] for o in opts], [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        iso = d.get("isolated", {}).get("value", 0)
        print(f"{wl:9s} {str(opts):22s} value {d['value']:.4g} us {1e3*d['ms_per_step']:.2f} isolated {iso:.4g} e2e {d['e2e']['value']:.4g} sync {d['e2e'].get('synchronous',{}).get('value',0):.4g} parity {d['parity_max_abs_err']} launches {d['gpu_launches']}", flush=True)
    except Exception as e:
        print(wl, opts, "FAILED", r.stdout[-300:], r.stderr[-1500:], flush=True)
for o in ("1", "0", "64", "56"):
    run("h2_sweep", ["frame32s=" + o])
run("h2_sweep", ["frame32s=1"], 10)
run("h2_sweep", ["frame32s=1"], 500)
run("h2_bonds", [], 20)
run("h2_qem", [], 20)
run("swap7_qem", [], 5)
PY
timeout 900 python bench.py > gpurun_out/r2_bench_v4.json 2> gpurun_out/r2_bench_v4.err; echo bench_rc=$?; tail -3 gpurun_out/r2_bench_v4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_v4.json'))
print('H2 value %.4g isolated %.4g e2e %.4g sync %.4g ms %.5f parity %s cpu %.4g' % (d['value'], d['isolated']['value'], d['e2e']['value'], d['e2e'].get('synchronous',{}).get('value',0), d['ms_per_step'], d['parity_max_abs_err'], d['cpu_baseline']['value']))
for k,v in d.get('also',{}).items():
    if 'error' in v: print(k, v); continue
    print(k, 'value %.5g ms %.4f iso %s e2e %.4g parity %s frac %.3f fexec %s cpu %s wall %s' % (v['value'], v['ms_per_step'], (v.get('isolated') or {}).get('value'), v['e2e']['value'], v['parity_max_abs_err'], v['roofline']['frac'], v['roofline'].get('frac_executed'), (v.get('cpu_baseline') or {}).get('value'), v['wall_s']))
PY

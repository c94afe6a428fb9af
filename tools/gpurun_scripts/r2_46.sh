set -x
timeout 600 python -m pytest tests/test_trajectories.py -q -m gpu -x --durations=5 2>&1 | tail -15
for o in 512 1024; do
  timeout 300 python bench.py --workload swap7_qem --no-also --no-cpu-baseline --steps 5 --warmup 3 --opt tile_threads=$o > gpurun_out/x_tile$o.json 2> gpurun_out/x_tile$o.err; echo rc=$?
  python - <<PY
import json
d=json.load(open('gpurun_out/x_tile$o.json'))
print('tile_threads=$o value %.5g ms %.4f parity %s iso %.5g est %s' % (d['value'], d['ms_per_step'], d['parity_max_abs_err'], d['isolated']['value'], d.get('estimator',{}).get('value')))
PY
done

# round 2, session o: split-phase host steps (two env groups), host steps launched as overlap mode 0
exec > gpurun_out/session_r2o.log 2>&1
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err
tail -2 gpurun_out/bench_r2o.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2o.json'))
print(d['value']/1e9, d['ms_per_step']*1e3, d['roofline']['frac'])
print(json.dumps(d['e2e'], indent=1))
for k,v in d['configs'].items(): print(k, v['ms_per_step']*1e3, v['roofline']['frac'], v['e2e']['value']/1e6, v['e2e']['ms_per_step']*1e3, v['e2e']['frac_of_ceiling'])
PY

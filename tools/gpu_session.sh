exec > gpurun_out/session.log 2>&1
set -x
RS_PER_MATCH=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
Q="--steps 4000 --warmup 800 --min-warmup 300 --cpu-seconds 0.3 --e2e-steps 10"
P='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("RESULT %.2f us  frac %.4f  e2e %.3g" % (d["ms_per_step"]*1000, d["roofline"]["frac"], d["e2e"]["value"]))'
for cfg in "RS_PER_MATCH=1" "RS_PER_MATCH=0" ; do
  echo "== $cfg"; env $cfg python bench.py $Q 2>&1 | python -c "$P"
done
for n in 4096 16384 131072 262144; do for cfg in "RS_PER_MATCH=1" "RS_PER_MATCH=0"; do
  echo "== envs $n $cfg"; env $cfg python bench.py $Q --envs $n 2>&1 | python -c "$P"
done; done
RS_PER_MATCH=1 ncu --set full --clock-control none --import-source on -k regex:k_vss_env_step --launch-skip 4810 --launch-count 2 -o gpurun_out/prof_pm3 python bench.py --steps 16 --warmup 4800 --no-graph --cpu-seconds 0.2 --e2e-steps 10 > gpurun_out/ncu_pm3.log 2>&1
tail -2 gpurun_out/ncu_pm3.log | cut -c1-200

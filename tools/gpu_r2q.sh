# round 2, session q (8 GPUs): the scaling bench as the driver runs it, final build; sharded == unsharded check
exec > gpurun_out/session_r2q.log 2>&1
set -x
nvidia-smi topo -m | head -14
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_r2q_n8.json 2> gpurun_out/bench_r2q_n8.err
tail -3 gpurun_out/bench_r2q_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 --cpu-seconds 2 --no-extras > gpurun_out/bench_r2q_n4.json 2> gpurun_out/bench_r2q_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-seconds 2 --no-extras > gpurun_out/bench_r2q_n2.json 2> gpurun_out/bench_r2q_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/multi_gpu_check.py 2>&1 | tail -3
python - <<'PY'
import json
for n in (8, 4, 2):
    try:
        d = json.load(open('gpurun_out/bench_r2q_n%d.json' % n))
    except Exception as e:
        print(n, 'no line', e); continue
    e = d['e2e']
    print(n, 'value %.2f G  %.2f us  e2e %.0f M (ceiling frac %.2f)  pipelined %s' % (d['value'] / 1e9, d['ms_per_step'] * 1e3, e['value'] / 1e6, e['frac_of_ceiling'],
          ('%.0f M' % (e['pipelined']['value'] / 1e6)) if 'pipelined' in e else '-'))
    for k in ('strong', 'gather'):
        if k in d: print('  ', k, json.dumps(d[k])[:300])
PY

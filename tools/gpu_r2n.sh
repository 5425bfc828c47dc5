# round 2, session n: host-step action staging by row size; GPU suite, e2e probe, default bench line
exec > gpurun_out/session_r2n.log 2>&1
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/e2e_probe.py
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err
tail -2 gpurun_out/bench_r2n.err

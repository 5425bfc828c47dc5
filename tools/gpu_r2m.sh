# round 2, session m: the driver's sequence (smoke, tests, reference arm, bench), then --config lines
exec > gpurun_out/session_r2m.log 2>&1
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2m_ref.json 2> gpurun_out/bench_r2m_ref.err
cut -c1-200 gpurun_out/bench_r2m_ref.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err
tail -2 gpurun_out/bench_r2m.err
for c in sd4096 cp16384 vss4096; do
  timeout 300 python bench.py --config $c --no-extras --steps 20 --warmup 5 --cpu-seconds 3 > gpurun_out/bench_r2m_$c.json 2> gpurun_out/bench_r2m_$c.err
  tail -2 gpurun_out/bench_r2m_$c.err
done
timeout 300 python tools/launch_floor.py

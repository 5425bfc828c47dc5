# round 2, session y: the driver's sequence on the final tree (smoke, GPU suite, reference arm, bench), --config lines, launch list
exec > gpurun_out/session_r2y.log 2>&1
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2y_ref.json 2> gpurun_out/bench_r2y_ref.err
cut -c1-200 gpurun_out/bench_r2y_ref.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err
tail -2 gpurun_out/bench_r2y.err
for c in sd4096 cp16384 vss4096; do
  timeout 300 python bench.py --config $c --no-extras --steps 20 --warmup 5 --cpu-seconds 3 --sustained-seconds 0 > gpurun_out/bench_r2y_$c.json 2> gpurun_out/bench_r2y_$c.err
  tail -2 gpurun_out/bench_r2y_$c.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2y.csv python bench.py --steps 20 --warmup 5 --min-ms 2 --cpu-seconds 0.2 --e2e-steps 10 --no-extras --sustained-seconds 0 > gpurun_out/launches_r2y.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2y.csv | tail -8

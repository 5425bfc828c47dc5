# round 2, session p: the driver's sequence on the final build (smoke, tests, reference arm, bench), --config lines,
# ncu launch list of the bench command, full captures of the headline kernel (dense build in rotation, serial build)
exec > gpurun_out/session_r2p.log 2>&1
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2p_ref.json 2> gpurun_out/bench_r2p_ref.err
cut -c1-200 gpurun_out/bench_r2p_ref.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err
tail -2 gpurun_out/bench_r2p.err
for c in sd4096 cp16384 vss4096; do
  timeout 300 python bench.py --config $c --no-extras --steps 20 --warmup 5 --cpu-seconds 3 > gpurun_out/bench_r2p_$c.json 2> gpurun_out/bench_r2p_$c.err
  tail -2 gpurun_out/bench_r2p_$c.err
done
timeout 300 python tools/launch_floor.py
timeout 300 python tools/e2e_probe.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2p.csv python bench.py --steps 20 --warmup 5 --min-ms 2 --cpu-seconds 0.2 --e2e-steps 10 --no-extras > gpurun_out/launches_r2p.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2p.csv | tail -15
N="timeout 600 ncu --set full --clock-control none --import-source on --launch-count 2"
RS_PER_MATCH=1 RS_STEP_OVERLAP=3 $N -k regex:k_vss_env_step --launch-skip 1210 -o gpurun_out/prof_r2p_vss65536_dense python tools/step_timing.py --task vss --envs 65536 --worlds 4 --warmup 300 --no-graph --steps 16 > gpurun_out/ncu_r2p_a.log 2>&1
RS_PER_MATCH=1 RS_STEP_OVERLAP=0 $N -k regex:k_vss_env_step --launch-skip 1210 -o gpurun_out/prof_r2p_vss65536_serial python tools/step_timing.py --task vss --envs 65536 --worlds 4 --warmup 300 --no-graph --steps 16 > gpurun_out/ncu_r2p_b.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4

exec > gpurun_out/session.log 2>&1
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
cat gpurun_out/bench_r1b.json | cut -c1-400
python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_r1b_ref.json 2>&1
# launch list of the same bench command (short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 64 --warmup 8 --min-warmup 8 --cpu-seconds 0.2 --e2e-steps 10 > gpurun_out/launches_r1b.log 2>&1
# full captures: headline kernel in steady state; lane-per-body VSS at 4096; SSL tasks at their config sizes
N="ncu --set full --clock-control none --import-source on --launch-count 2"
$N -k regex:k_vss_env_step --launch-skip 4810 -o gpurun_out/prof_r1b_vss65536 python bench.py --steps 16 --warmup 4800 --no-graph --cpu-seconds 0.2 --e2e-steps 10 > gpurun_out/ncu_a.log 2>&1
$N -k regex:k_vss_env_step --launch-skip 2410 -o gpurun_out/prof_r1b_vss4096 python tools/step_timing.py --task vss --envs 4096 --no-graph --steps 16 > gpurun_out/ncu_b.log 2>&1
$N -k regex:k_ssl_env_step --launch-skip 2410 -o gpurun_out/prof_r1b_sd4096 python tools/step_timing.py --task sd --envs 4096 --no-graph --steps 16 > gpurun_out/ncu_c.log 2>&1
$N -k regex:k_ssl_env_step --launch-skip 2410 -o gpurun_out/prof_r1b_cp16384 python tools/step_timing.py --task cp --envs 16384 --no-graph --steps 16 > gpurun_out/ncu_d.log 2>&1
T="python tools/step_timing.py"
$T --task vss --envs 4096; $T --task sd --envs 4096; $T --task cp --envs 16384; $T --task vss --envs 32768; $T --task vss --envs 65536

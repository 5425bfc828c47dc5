# one GPU session that regenerates every measured artefact of a round: tests, bench line,
# reference arm, launch list, full ncu captures, per-config timings.  usage (on the GPU box):
#   bash tools/gpu_session.sh r1c
R=${1:-r1c}
exec > gpurun_out/session_$R.log 2>&1
set -x
python -c "import __graft_entry__ as g; g.smoke()"
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
cut -c1-400 gpurun_out/bench_$R.json
python bench.py --impl reference --steps 50 --warmup 5 > gpurun_out/bench_${R}_ref.json 2>&1
# launch list of the same bench command (short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 64 --warmup 8 --min-warmup 8 --cpu-seconds 0.2 --e2e-steps 10 > gpurun_out/launches_$R.log 2>&1
# full captures: headline kernel in steady state; SSL tasks at their config sizes
N="ncu --set full --clock-control none --import-source on --launch-count 2"
$N -k regex:k_vss_env_step --launch-skip 4810 -o gpurun_out/prof_${R}_vss65536 python bench.py --steps 16 --warmup 4800 --no-graph --cpu-seconds 0.2 --e2e-steps 10 > gpurun_out/ncu_a.log 2>&1
$N -k regex:k_vss_env_step --launch-skip 2410 -o gpurun_out/prof_${R}_vss4096 python tools/step_timing.py --task vss --envs 4096 --no-graph --steps 16 > gpurun_out/ncu_b.log 2>&1
$N -k regex:k_ssl_env_step --launch-skip 2410 -o gpurun_out/prof_${R}_sd4096 python tools/step_timing.py --task sd --envs 4096 --no-graph --steps 16 > gpurun_out/ncu_c.log 2>&1
$N -k regex:k_ssl_env_step --launch-skip 2410 -o gpurun_out/prof_${R}_cp16384 python tools/step_timing.py --task cp --envs 16384 --no-graph --steps 16 > gpurun_out/ncu_d.log 2>&1
T="python tools/step_timing.py"
$T --task vss --envs 4096; $T --task sd --envs 4096; $T --task cp --envs 16384; $T --task vss --envs 32768; $T --task vss --envs 65536
$T --task vss --envs 131072; $T --task vss --envs 262144; $T --task vss --envs 1048576 --worlds 2 --steps 400 --warmup 100

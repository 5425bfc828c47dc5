# round 2, session j: compute-sanitizer on every step kernel (overlap modes included), ncu launch list of the bench command, SSL capture
exec > gpurun_out/session_r2j.log 2>&1
set -x
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py 2>&1 | tail -4
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_run.py 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2j.csv python bench.py --steps 20 --warmup 5 --min-ms 2 --cpu-seconds 0.2 --e2e-steps 10 --no-extras > gpurun_out/launches_r2j.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r2j.csv | tail -15
N="timeout 600 ncu --set full --clock-control none --import-source on --launch-count 2"
RS_PER_MATCH=1 RS_STEP_OVERLAP=0 $N -k regex:k_ssl_env_step --launch-skip 1210 -o gpurun_out/prof_r2j_sd65536 python tools/step_timing.py --task sd --envs 65536 --worlds 4 --warmup 300 --no-graph --steps 16 > gpurun_out/ncu_r2j_a.log 2>&1
RS_PER_MATCH=1 RS_STEP_OVERLAP=3 $N -k regex:k_vss_env_step --launch-skip 1210 -o gpurun_out/prof_r2j_vss65536_dense python tools/step_timing.py --task vss --envs 65536 --worlds 4 --warmup 300 --no-graph --steps 16 > gpurun_out/ncu_r2j_b.log 2>&1
RS_PER_MATCH=1 RS_STEP_OVERLAP=3 $N -k regex:k_vss_env_step --launch-skip 604 -o gpurun_out/prof_r2j_vss1m_dense python tools/step_timing.py --task vss --envs 1048576 --worlds 2 --warmup 300 --no-graph --steps 8 > gpurun_out/ncu_r2j_c.log 2>&1

# round 2, session z: full captures of the headline kernel at 1 M matches on the final build (the multi-wave regime of the overlapped bench region)
exec > gpurun_out/session_r2z.log 2>&1
set -x
N="timeout 600 ncu --set full --clock-control none --import-source on --launch-count 2"
RS_PER_MATCH=1 RS_STEP_OVERLAP=3 $N -k regex:k_vss_env_step --launch-skip 604 -o gpurun_out/prof_r2z_vss1m_dense python tools/step_timing.py --task vss --envs 1048576 --worlds 2 --warmup 300 --no-graph --steps 8 > gpurun_out/ncu_r2z_c.log 2>&1
RS_PER_MATCH=1 RS_STEP_OVERLAP=0 $N -k regex:k_vss_env_step --launch-skip 604 -o gpurun_out/prof_r2z_vss1m_serial python tools/step_timing.py --task vss --envs 1048576 --worlds 2 --warmup 300 --no-graph --steps 8 > gpurun_out/ncu_r2z_d.log 2>&1
ls -la gpurun_out/prof_r2z*

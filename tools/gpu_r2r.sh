# round 2, session r: full ncu captures of the kernels of BASELINE configs 2, 3, 4 at their sizes (final build), serialised
exec > gpurun_out/session_r2r.log 2>&1
set -x
N="timeout 600 ncu --set full --clock-control none --import-source on --launch-count 2"
$N -k regex:k_vss_env_step --launch-skip 2410 -o gpurun_out/prof_r2r_vss4096 python tools/step_timing.py --task vss --envs 4096 --no-graph --steps 16 > gpurun_out/ncu_r2r_b.log 2>&1
$N -k regex:k_ssl_env_step --launch-skip 2410 -o gpurun_out/prof_r2r_sd4096 python tools/step_timing.py --task sd --envs 4096 --no-graph --steps 16 > gpurun_out/ncu_r2r_c.log 2>&1
$N -k regex:k_ssl_env_step --launch-skip 2410 -o gpurun_out/prof_r2r_cp16384 python tools/step_timing.py --task cp --envs 16384 --no-graph --steps 16 > gpurun_out/ncu_r2r_d.log 2>&1
T="python tools/step_timing.py"
for ov in 0 3; do
RS_STEP_OVERLAP=$ov $T --task vss --envs 4096 --worlds 125; RS_STEP_OVERLAP=$ov $T --task sd --envs 4096 --worlds 133; RS_STEP_OVERLAP=$ov $T --task cp --envs 16384 --worlds 67
RS_STEP_OVERLAP=$ov $T --task vss --envs 32768 --worlds 16; RS_STEP_OVERLAP=$ov $T --task vss --envs 65536
RS_STEP_OVERLAP=$ov $T --task vss --envs 262144 --worlds 2 --steps 1000; RS_STEP_OVERLAP=$ov $T --task vss --envs 1048576 --worlds 2 --steps 400 --warmup 100
done
ls -la gpurun_out/prof_r2r*

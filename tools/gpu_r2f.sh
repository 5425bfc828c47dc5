# round 2, session f: SSL kernels after the shared-memory contact resolve
exec > gpurun_out/session_r2f.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
T="timeout 300 python tools/step_timing.py --steps 6000"
for ov in 0 3; do
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task sd --envs 65536 --worlds 8 | sed "s/^/ov=$ov worlds=8 /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task sd --envs 16384 --worlds 32 | sed "s/^/ov=$ov /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task sd --envs 4096 --worlds 133 | sed "s/^/ov=$ov /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task cp --envs 16384 --worlds 67 | sed "s/^/ov=$ov /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task step_ssl --envs 65536 --worlds 8 | sed "s/^/ov=$ov /"
  RS_PER_MATCH=0 RS_STEP_OVERLAP=$ov $T --task step_ssl --envs 65536 --worlds 8 | sed "s/^/ov=$ov /"
done
RS_STEP_OVERLAP=0 $T --task sd --envs 4096 --worlds 133 | sed "s/^/auto ov=0 /"
RS_STEP_OVERLAP=0 $T --task sd --envs 16384 --worlds 32 | sed "s/^/auto ov=0 /"
N="timeout 600 ncu --set full --clock-control none --import-source on --launch-count 2"
RS_PER_MATCH=1 RS_STEP_OVERLAP=0 $N -k regex:k_ssl_env_step --launch-skip 1210 -o gpurun_out/prof_r2f_sd65536 python tools/step_timing.py --task sd --envs 65536 --worlds 4 --warmup 300 --no-graph --steps 16 > gpurun_out/ncu_r2f_a.log 2>&1

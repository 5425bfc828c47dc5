# round 2, session l: SSL out-of-line goal walls
exec > gpurun_out/session_r2l.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
T="timeout 300 python tools/step_timing.py --steps 6000"
for ov in 0 3; do
RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task sd --envs 65536 --worlds 8 | sed "s/^/ov=$ov /"
RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task sd --envs 4096 --worlds 133 | sed "s/^/ov=$ov /"
RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task cp --envs 16384 --worlds 67 | sed "s/^/ov=$ov /"
RS_PER_MATCH=0 RS_STEP_OVERLAP=$ov $T --task sd --envs 4096 --worlds 133 | sed "s/^/ov=$ov /"
RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task step_ssl --envs 65536 --worlds 8 | sed "s/^/ov=$ov /"
done

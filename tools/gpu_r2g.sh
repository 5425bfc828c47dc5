# round 2, session g: release-store hand-over, SSL info-by-RED + lean walls
exec > gpurun_out/session_r2g.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
T="timeout 300 python tools/step_timing.py --steps 6000"
for ov in 0 3; do
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task sd --envs 65536 --worlds 8 | sed "s/^/ov=$ov worlds=8 /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task sd --envs 4096 --worlds 133 | sed "s/^/ov=$ov /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task cp --envs 16384 --worlds 67 | sed "s/^/ov=$ov /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task cp --envs 65536 --worlds 16 | sed "s/^/ov=$ov /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task vss --envs 65536 --worlds 8 | sed "s/^/ov=$ov /"
done
timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_r2g_full.json 2> gpurun_out/bench_r2g_full.err
tail -3 gpurun_out/bench_r2g_full.err

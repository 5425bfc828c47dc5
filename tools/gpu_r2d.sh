# round 2, session d: tests (all kernels under the overlap protocol), the full default bench line, mapping x overlap matrix
exec > gpurun_out/session_r2d.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^Environment init" | tail -40
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2d_full.json 2> gpurun_out/bench_r2d_full.err
tail -3 gpurun_out/bench_r2d_full.err
T="timeout 300 python tools/step_timing.py --steps 6000"
for pm in 0 1; do for ov in 0 2 3; do
  RS_PER_MATCH=$pm RS_STEP_OVERLAP=$ov $T --task vss --envs 4096 --worlds 128 | sed "s/^/ov=$ov /"
  RS_PER_MATCH=$pm RS_STEP_OVERLAP=$ov $T --task sd --envs 4096 --worlds 133 | sed "s/^/ov=$ov /"
  RS_PER_MATCH=$pm RS_STEP_OVERLAP=$ov $T --task cp --envs 16384 --worlds 67 | sed "s/^/ov=$ov /"
done; done
for ov in 0 2 3; do
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task vss --envs 65536 --worlds 1 | sed "s/^/ov=$ov worlds=1 /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task vss --envs 65536 --worlds 8 | sed "s/^/ov=$ov worlds=8 /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task sd --envs 65536 --worlds 8 | sed "s/^/ov=$ov worlds=8 /"
  RS_PER_MATCH=0 RS_STEP_OVERLAP=$ov $T --task vss --envs 4096 --worlds 1 | sed "s/^/ov=$ov worlds=1 /"
  RS_PER_MATCH=1 RS_STEP_OVERLAP=$ov $T --task vss --envs 4096 --worlds 1 | sed "s/^/ov=$ov worlds=1 /"
done

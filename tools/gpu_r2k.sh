# round 2, session k: lock merged into the step counter
exec > gpurun_out/session_r2k.log 2>&1
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_r2k_full.json 2> gpurun_out/bench_r2k_full.err
tail -3 gpurun_out/bench_r2k_full.err
T="timeout 300 python tools/step_timing.py --steps 6000"
RS_PER_MATCH=1 RS_STEP_OVERLAP=3 $T --task sd --envs 65536 --worlds 8
RS_PER_MATCH=0 RS_STEP_OVERLAP=2 $T --task sd --envs 4096 --worlds 133
RS_PER_MATCH=0 RS_STEP_OVERLAP=0 $T --task sd --envs 4096 --worlds 133

# round 2, session e: load-spin acquire, per-match kernels in rotation, denser builds
exec > gpurun_out/session_r2e.log 2>&1
set -x
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "overlap or options or rs_step_between" 2>&1 | tail -5
rm -f gpurun_out/variants.txt
timeout 600 python tools/variants.py run --task vss --sizes 32768,65536 --mode 1 --steps 8000 --worlds 8 --env RS_STEP_OVERLAP=3
timeout 300 python tools/variants.py run --task vss --sizes 65536 --mode 1 --steps 8000 --worlds 1 --env RS_STEP_OVERLAP=2 --only base
timeout 300 python tools/variants.py run --task vss --sizes 65536 --mode 1 --steps 8000 --worlds 1 --env RS_STEP_OVERLAP=0 --only base
timeout 900 python bench.py --steps 20 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_r2e_full.json 2> gpurun_out/bench_r2e_full.err
tail -3 gpurun_out/bench_r2e_full.err
timeout 900 python bench.py --steps 240 --warmup 8 --cpu-seconds 2 --no-extras > gpurun_out/bench_r2e_k240.json 2> gpurun_out/bench_r2e_k240.err

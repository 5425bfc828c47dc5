mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/run13.log
for n in 4096 65536; do RS_PER_MATCH=1 python tools/step_timing.py --task vss --envs $n >> gpurun_out/run13.log 2>&1; done
cat gpurun_out/run13.log

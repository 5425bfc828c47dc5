mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/run8.log
for n in 4096 8192 12288 16384 20480 24576 32768; do
  RS_PER_MATCH=1 python tools/step_timing.py --task vss --envs $n >> gpurun_out/run8.log 2>&1
  RS_PER_MATCH=0 python tools/step_timing.py --task vss --envs $n >> gpurun_out/run8.log 2>&1
done
cat gpurun_out/run8.log

mkdir -p gpurun_out; rm -f gpurun_out/run14.log
for g in 8 32 128; do RS_PER_MATCH=1 python tools/step_timing.py --task vss --envs 65536 --graph-steps $g --steps 8192 >> gpurun_out/run14.log 2>&1; done
RS_PER_MATCH=1 python tools/step_timing.py --task vss --envs 65536 --no-graph --steps 8192 >> gpurun_out/run14.log 2>&1
cat gpurun_out/run14.log

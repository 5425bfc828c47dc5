# round 2, session b: full GPU tests, occupancy variants under step overlap, single-world chaining, ncu launch list
exec > gpurun_out/session_r2b.log 2>&1
set -x
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^Environment init" | tail -60
rm -f gpurun_out/variants.txt
for ov in 0 2; do
  timeout 600 python tools/variants.py run --task vss --sizes 32768,65536 --mode 1 --steps 6000 --worlds 8 --env RS_STEP_OVERLAP=$ov
done
timeout 300 python tools/variants.py run --task vss --sizes 65536 --mode 1 --steps 6000 --worlds 1 --env RS_STEP_OVERLAP=0 --only base,t576
timeout 300 python tools/variants.py run --task vss --sizes 65536 --mode 1 --steps 6000 --worlds 1 --env RS_STEP_OVERLAP=2 --only base,t576
timeout 300 python tools/variants.py run --task vss --sizes 65536 --mode 1 --steps 6000 --worlds 8 --env RS_STEP_OVERLAP=2,RS_BLOCK=32 --only base,t576
timeout 300 python tools/variants.py run --task vss --sizes 65536,1048576 --mode 1 --steps 2000 --worlds 2 --env RS_STEP_OVERLAP=2 --only base,t576,t640
cat gpurun_out/variants.txt

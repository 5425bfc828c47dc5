mkdir -p gpurun_out; rm -f gpurun_out/variants.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/run3.log
python tools/variants.py run --task vss --sizes 4096,65536,262144 --mode 1 >> gpurun_out/run3.log 2>&1
python tools/variants.py run --task vss --sizes 4096,65536,262144 --mode 1 --worlds 1 --only base,smem.,smem_sub0 >> gpurun_out/run3.log 2>&1
python tools/variants.py run --task vss --sizes 65536 --mode 1 --env RS_BLOCK=32 --only smem. >> gpurun_out/run3.log 2>&1
python tools/variants.py run --task vss --sizes 65536 --mode 1 --env RS_BLOCK=128 --only smem. >> gpurun_out/run3.log 2>&1
python tools/variants.py run --task vss --sizes 65536 --mode 1 --env RS_PDL=0 --only smem. >> gpurun_out/run3.log 2>&1
tail -40 gpurun_out/run3.log

mkdir -p gpurun_out; rm -f gpurun_out/variants.txt
python tools/variants.py run --task vss --sizes 4096,16384,65536,262144 --mode 1 > gpurun_out/run18.log 2>&1
RS_LIB=build/variants/lib_cur2.so timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/run18.log
cat gpurun_out/run18.log

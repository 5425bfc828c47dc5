mkdir -p gpurun_out; rm -f gpurun_out/run17.log
for v in "" "RS_HOST_ZEROCOPY=1" "RS_HOST_CHUNKED=1"; do
env $v python bench.py --steps 2400 --warmup 800 --min-warmup 100 --cpu-seconds 0.5 --e2e-steps 200 > gpurun_out/bench_tmp.json 2>> gpurun_out/run17.log
python -c "import json; d=json.load(open('gpurun_out/bench_tmp.json')); print('$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pcie_gbs'])" >> gpurun_out/run17.log
done
RS_HOST_ZEROCOPY=1 timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/run17.log
cat gpurun_out/run17.log

# round 2, session h (2 GPUs): device guard test, sharded == unsharded over NCCL, bench --gpus 2, launch floor
exec > gpurun_out/session_r2h.log 2>&1
set -x
nvidia-smi topo -m
timeout 300 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "own_device" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 2>&1 | tail -8
timeout 300 python tools/launch_floor.py
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_r2h_n2.json 2> gpurun_out/bench_r2h_n2.err
tail -3 gpurun_out/bench_r2h_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2h_n2_ref.json 2> gpurun_out/bench_r2h_n2_ref.err
cut -c1-300 gpurun_out/bench_r2h_n2_ref.json

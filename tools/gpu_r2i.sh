# round 2, session i (8 GPUs): the scaling bench as the driver runs it
exec > gpurun_out/session_r2i.log 2>&1
set -x
nvidia-smi topo -m | head -14
numactl -H 2>/dev/null | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --cpu-seconds 2 > gpurun_out/bench_r2i_n8.json 2> gpurun_out/bench_r2i_n8.err
tail -3 gpurun_out/bench_r2i_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 --cpu-seconds 2 --no-extras > gpurun_out/bench_r2i_n4.json 2> gpurun_out/bench_r2i_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/multi_gpu_check.py 2>&1 | tail -3

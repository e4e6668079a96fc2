#!/bin/bash
# 8-GPU weak-scaling point of bench.py on one box (launched exactly as the driver does)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/${TAG}_scale_err_8.log | tee gpurun_out/${TAG}_scale_8.json | cut -c1-330
tail -2 gpurun_out/${TAG}_scale_err_8.log | cut -c1-200

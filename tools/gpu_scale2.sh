#!/bin/bash
# 2-GPU check of the data-parallel bench path (launched exactly as the driver does) + the N=1 line on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/${TAG}_scale_err_2.log | tee gpurun_out/${TAG}_scale_2.json | cut -c1-420
tail -3 gpurun_out/${TAG}_scale_err_2.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/${TAG}_scale_1.json | cut -c1-420
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | cut -c1-300

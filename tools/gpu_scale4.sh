#!/bin/bash
# 1 / 2 / 4 GPU weak-scaling check of bench.py on one box (launched exactly as the driver does)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi -L | head -8
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 2>gpurun_out/${TAG}_scale_err_$n.log | tee gpurun_out/${TAG}_scale_$n.json | cut -c1-330
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/${TAG}_scale_1.json | cut -c1-330

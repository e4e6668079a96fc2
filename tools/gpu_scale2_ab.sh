#!/bin/bash
# 2-GPU box: NCCL correctness of the overlapped exchange, then bench.py at N=2 with and without the overlap
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02w}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/ddp_check.py 2>gpurun_out/${TAG}_ddp_check.err | tee gpurun_out/${TAG}_ddp_check.log
echo "ddp_check rc=$?"; tail -3 gpurun_out/${TAG}_ddp_check.err | cut -c1-300
for mode in 1 0; do
  B200UNET_OVERLAP_ALLREDUCE=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$mode bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/${TAG}_scale2_overlap${mode}.err | tee -a gpurun_out/${TAG}_scale2_overlap${mode}.json | cut -c1-200
  tail -2 gpurun_out/${TAG}_scale2_overlap${mode}.err | cut -c1-200
done

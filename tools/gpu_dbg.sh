#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02k}
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_northstar.py tests/test_gpu_dynunet.py -q -x 2>&1 | tail -8
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -1 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log
for a in "32 32 128 res" "32 64 128 mode1"; do
  echo "## $a" >> gpurun_out/${TAG}_halo_timeline.txt
  timeout 120 python tools/halo_timeline.py $a >> gpurun_out/${TAG}_halo_timeline.txt 2>&1
done
grep "steady-state\|epilogue warp 2" gpurun_out/${TAG}_halo_timeline.txt
B200UNET_HALO_SIDE_RING=0 timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times_noring.csv > gpurun_out/${TAG}_layer_times_noring.log 2>&1; head -1 gpurun_out/${TAG}_layer_times_noring.log; tail -1 gpurun_out/${TAG}_layer_times_noring.log

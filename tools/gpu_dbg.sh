#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02v}
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_dynunet.py tests/test_gpu_northstar.py -q -x 2>&1 | tail -3
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -1 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log
grep "dgrad encoder.downsampling" gpurun_out/${TAG}_layer_times.csv
B200UNET_CLASS_PAIR=0 timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times_nopair.csv > gpurun_out/${TAG}_layer_times_nopair.log 2>&1; head -1 gpurun_out/${TAG}_layer_times_nopair.log
grep "dgrad encoder.downsampling" gpurun_out/${TAG}_layer_times_nopair.csv

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02n}
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_northstar.py tests/test_gpu_dynunet.py -q -x 2>&1 | tail -3
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -1 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log
for a in "8 32 128 plain" "32 32 128 plain"; do
  echo "## $a" >> gpurun_out/${TAG}_halo_timeline.txt
  timeout 120 python tools/halo_timeline.py $a >> gpurun_out/${TAG}_halo_timeline.txt 2>&1
done
grep "##\|steady-state\|kernel time\|deltas of tiles" gpurun_out/${TAG}_halo_timeline.txt

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "--- MAXC=256 BN=64"
( B200UNET_HALO_MAXC=256 B200UNET_HALO_BN=64 timeout 600 python tools/conv_bench.py epi 10 2>&1 | grep convbench ) | grep -E '"ci": (128|256)' | cut -c1-250
( B200UNET_HALO_MAXC=256 B200UNET_HALO_BN=64 timeout 600 python tools/conv_bench.py fwd 10 2>&1 | grep convbench ) | grep -E '"ci": (128|256)' | cut -c1-250
timeout 300 python tools/layer_times.py gpurun_out/layer_times13.csv 2>&1 | tail -1 | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "groupnorm" 2>&1 | tail -2

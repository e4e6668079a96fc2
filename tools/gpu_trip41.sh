#!/bin/bash
cd "$(dirname "$0")/.."
echo "--- WGHALO_MAXC=256"
( B200UNET_WGHALO_MAXC=256 timeout 600 python tools/conv_bench.py wgrad 10 2>&1 | grep convbench ) | grep -E '"ci": (128|256)|total' | cut -c1-250
B200UNET_WGHALO_MAXC=256 timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "weight_gradient" 2>&1 | tail -2

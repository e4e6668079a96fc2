#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "weight_gradient or groupnorm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
( timeout 600 python tools/conv_bench.py wgrad 10 2>&1 | grep convbench ) | cut -c1-200
timeout 300 python tools/layer_times.py gpurun_out/layer_times13.csv 2>&1 | tail -1 | cut -c1-400

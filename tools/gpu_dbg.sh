#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02d}
timeout 900 python tools/graph_debug.py > gpurun_out/${TAG}_graph_debug.log 2>&1; cat gpurun_out/${TAG}_graph_debug.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_dynunet.py tests/test_gpu_model.py -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -1 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log

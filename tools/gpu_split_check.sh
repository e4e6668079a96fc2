#!/bin/bash
# two-part backward / split-graph step: parity tests on one GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02w}
timeout 900 python -m pytest tests/test_gpu_prepost.py tests/test_gpu_dynunet.py -q -k "two_part or graphed or epoch_training or deterministic or dynunet" > gpurun_out/${TAG}_split_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_split_pytest.log

#!/bin/bash
cd "$(dirname "$0")/.."
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "linear" 2>&1 | grep -E "^E|assert|passed|failed" | head -12; done

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 python tools/run_configs.py 2>&1 | grep -v Warn | tail -3 | tee gpurun_out/configs_c3_c5.json

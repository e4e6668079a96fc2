#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/umma_rate.py 2>&1 | grep -v Warn | tee gpurun_out/umma_rate.txt | sed -n 15,36p

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/umma_rate.py 2>&1 | grep -v Warn | tee gpurun_out/umma_rate.txt | head -48
B200UNET_HALO_MAXC=256 timeout 120 python tools/halo_timeline.py 128 128 64 plain 2>&1 | grep -v Warn | tail -22


#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02i}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/${TAG}_halo_mode1_32_64_128 python tools/halo_timeline.py 32 64 128 mode1 > gpurun_out/ncu_m1.log 2>&1; tail -2 gpurun_out/ncu_m1.log
ls -la gpurun_out/*.ncu-rep

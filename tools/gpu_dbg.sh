#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02s}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_igemm_conv -s 72 -c 4 -f -o gpurun_out/${TAG}_igemm_32_32 python tools/prof_step.py 2 > gpurun_out/ncu_cls.log 2>&1; tail -2 gpurun_out/ncu_cls.log
ls -la gpurun_out/${TAG}_igemm_32_32.ncu-rep

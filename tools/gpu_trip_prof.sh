#!/bin/bash
# round-end profiling: launch list of the bench command + --set full captures of the top kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log; echo
wc -l gpurun_out/r01_launches.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_halo -s 52 -c 8 -f -o gpurun_out/r01_conv_halo python tools/prof_step.py 2 > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wgrad_halo -s 12 -c 3 -f -o gpurun_out/r01_wgrad_halo python tools/prof_step.py 2 > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_igemm_conv|k_wgrad<" -s 40 -c 6 -f -o gpurun_out/r01_streaming python tools/prof_step.py 2 > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# round profiling: ncu launch list of the bench command (time + DRAM bytes per launch) and --set full captures of the top
# kernels inside a training step.  Numbers printed under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/${TAG}_bench_under_ncu.log; echo
wc -l gpurun_out/${TAG}_launches.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_halo -s 52 -c 8 -f -o gpurun_out/${TAG}_conv_halo python tools/prof_step.py 2 > gpurun_out/ncu_a.log 2>&1; tail -2 gpurun_out/ncu_a.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_wgrad_halo -s 12 -c 3 -f -o gpurun_out/${TAG}_wgrad_halo python tools/prof_step.py 2 > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_igemm_conv|k_wgrad<" -s 40 -c 6 -f -o gpurun_out/${TAG}_streaming python tools/prof_step.py 2 > gpurun_out/ncu_c.log 2>&1; tail -2 gpurun_out/ncu_c.log
python tools/summarize_ncu.py launches gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches.txt 2>&1 | tail -3
for k in conv_halo wgrad_halo streaming; do python tools/summarize_ncu.py full gpurun_out/${TAG}_$k.ncu-rep gpurun_out/${TAG}_${k}_ncu.txt 2>&1 | tail -1; done
ls -la gpurun_out/*.ncu-rep gpurun_out/${TAG}_*.txt

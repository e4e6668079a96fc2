#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/diag.log
echo "######## conv diag (halo + streaming)" | tee -a gpurun_out/diag.log
timeout 600 python tools/gpu_diag.py conv >> gpurun_out/diag.log 2>&1; echo "exit=$?" | tee -a gpurun_out/diag.log
grep -E "^\[diag\]|EXCEPTION|exit=" gpurun_out/diag.log | cut -c1-250 | tail -60
( timeout 600 python tools/conv_bench.py fwd 10 2>&1 | grep convbench ) | tee gpurun_out/convbench_halo.log
( B200UNET_NO_HALO=1 timeout 600 python tools/conv_bench.py fwd 10 2>&1 | grep convbench ) | tee gpurun_out/convbench_stream.log
( B200UNET_HALO_TD=2 timeout 600 python tools/conv_bench.py fwd 10 2>&1 | grep convbench ) | tee gpurun_out/convbench_td2.log
( B200UNET_HALO_BN=64 timeout 600 python tools/conv_bench.py fwd 10 2>&1 | grep convbench ) | tee gpurun_out/convbench_bn64.log
( timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -8 ) | tee gpurun_out/pytest_ops.log
( timeout 300 python tools/gpu_diag.py bench 2>&1 | grep diag ) | tee gpurun_out/bench_diag.log

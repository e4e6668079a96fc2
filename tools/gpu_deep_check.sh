#!/bin/bash
# deep pipeline ring of the streaming conv kernel for launches with <= one CTA per SM (B200UNET_IGEMM_DEEP=1): parity suite, bench
# line and per-launch times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02y}
export B200UNET_IGEMM_DEEP=1
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_deep.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_deep.log
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/${TAG}_bench_C2_deep.err > gpurun_out/${TAG}_bench_C2_deep.json; tail -2 gpurun_out/${TAG}_bench_C2_deep.err; cut -c1-330 gpurun_out/${TAG}_bench_C2_deep.json
timeout 300 python tools/layer_times.py gpurun_out/${TAG}_layer_times_deep.csv > gpurun_out/${TAG}_layer_times_deep.log 2>&1; head -1 gpurun_out/${TAG}_layer_times_deep.log; tail -1 gpurun_out/${TAG}_layer_times_deep.log
grep "16x16x16->256ch" gpurun_out/${TAG}_layer_times_deep.csv | head -4

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02l}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -1 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log
for a in "32 64 128 mode1" "32 8 128 mode1"; do
  echo "## $a" >> gpurun_out/${TAG}_halo_timeline.txt
  timeout 120 python tools/halo_timeline.py $a >> gpurun_out/${TAG}_halo_timeline.txt 2>&1
done
grep "steady-state\|epilogue warp 2" gpurun_out/${TAG}_halo_timeline.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench_C2.err; cut -c1-300 gpurun_out/${TAG}_bench_C2.json; tail -c 300 gpurun_out/${TAG}_bench_C2.err

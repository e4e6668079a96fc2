#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 600 python tools/graph_debug.py > gpurun_out/${TAG}_graph_debug.log 2>&1; cat gpurun_out/${TAG}_graph_debug.log
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${TAG}_pytest.log; tail -12 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -1 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log
for a in "32 32 128 plain" "32 32 128 res" "32 64 128 mode1" "64 32 128 plain"; do
  echo "## $a" >> gpurun_out/${TAG}_halo_timeline.txt
  timeout 120 python tools/halo_timeline.py $a >> gpurun_out/${TAG}_halo_timeline.txt 2>&1
done
grep "steady-state\|epilogue warp 2" gpurun_out/${TAG}_halo_timeline.txt
for a in "32 32 128 res" "32 64 128 mode1"; do
  tag=$(echo $a | tr ' ' '_')
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/${TAG}_halo_$tag python tools/halo_timeline.py $a > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log
done
for c in C2 C3 C5; do
  timeout 900 python bench.py --config $c 2>gpurun_out/${TAG}_bench_$c.err > gpurun_out/${TAG}_bench_$c.json
  tail -2 gpurun_out/${TAG}_bench_$c.err; cut -c1-700 gpurun_out/${TAG}_bench_$c.json
done

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02h}
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_gpu_northstar.py -q -x 2>&1 | tail -3
cat gpurun_out/northstar_measured.txt
for pf in 0 1; do
  B200UNET_HALO_SIDE_PF=$pf timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times_pf$pf.csv > gpurun_out/${TAG}_layer_times_pf$pf.log 2>&1; echo "side_pf=$pf"; head -1 gpurun_out/${TAG}_layer_times_pf$pf.log; tail -1 gpurun_out/${TAG}_layer_times_pf$pf.log
done
for a in "32 32 128 res" "32 64 128 mode1"; do
  echo "## $a" >> gpurun_out/${TAG}_halo_timeline.txt
  timeout 120 python tools/halo_timeline.py $a >> gpurun_out/${TAG}_halo_timeline.txt 2>&1
done
grep "steady-state\|epilogue warp 2" gpurun_out/${TAG}_halo_timeline.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench_C2.err; cut -c1-300 gpurun_out/${TAG}_bench_C2.json

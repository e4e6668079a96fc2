#!/bin/bash
# A/B measurements of the round-2 kernel changes (each selectable by an environment switch) + the two-issuer variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -3 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log
B200UNET_OLD_SMALL_OPS=1 B200UNET_S2_ZERO_INSERT=1 B200UNET_HALO_1X1_DENSE=0 timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times_r01paths.csv > gpurun_out/${TAG}_layer_times_r01paths.log 2>&1
head -1 gpurun_out/${TAG}_layer_times_r01paths.log; tail -1 gpurun_out/${TAG}_layer_times_r01paths.log
timeout 600 python tools/conv_bench.py all 10 > gpurun_out/${TAG}_convbench.log 2>&1; grep "weighted" gpurun_out/${TAG}_convbench.log
# compile-time variant: one (kh,kw) box per weight stage (the round-1 issue loop)
B200UNET_LIB=$PWD/3dunetcnn_b200/libb200unet_kws1.so timeout 600 python tools/conv_bench.py epi 10 > gpurun_out/${TAG}_convbench_kws1.log 2>&1; grep "weighted" gpurun_out/${TAG}_convbench_kws1.log
B200UNET_HALO_ISSUERS=2 timeout 600 python tools/conv_bench.py epi 10 > gpurun_out/${TAG}_convbench_ni2.log 2>&1; grep "weighted" gpurun_out/${TAG}_convbench_ni2.log
B200UNET_HALO_ISSUERS=2 timeout 600 python tools/conv_determinism.py 3 > gpurun_out/${TAG}_det_ni2.log 2>&1; tail -2 gpurun_out/${TAG}_det_ni2.log
B200UNET_HALO_ISSUERS=2 timeout 900 python -m pytest tests/test_gpu_ops.py -q -k "conv3d" 2>&1 | tail -3 > gpurun_out/${TAG}_ops_ni2.log; cat gpurun_out/${TAG}_ops_ni2.log
for a in "8 32 128 plain" "32 32 128 plain" "32 32 128 res" "64 32 128 plain" "32 64 128 mode1"; do
  echo "## $a" >> gpurun_out/${TAG}_halo_timeline.txt
  timeout 120 python tools/halo_timeline.py $a >> gpurun_out/${TAG}_halo_timeline.txt 2>&1
done
grep "steady-state" gpurun_out/${TAG}_halo_timeline.txt

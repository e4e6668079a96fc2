#!/bin/bash
# ncu --set full capture of the halo conv kernel for one shape/mode (source-level stall attribution)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200UNET_HALO_MAXC=256
for a in "128 128 64 plain" "32 32 128 plain"; do
  tag=$(echo $a | tr ' ' '_')
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_halo -s 2 -c 1 -f -o gpurun_out/halo_$tag python tools/halo_timeline.py $a > gpurun_out/ncu_halo_$tag.log 2>&1
  tail -3 gpurun_out/ncu_halo_$tag.log
done
ls -la gpurun_out/*.ncu-rep

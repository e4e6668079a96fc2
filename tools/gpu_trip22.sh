#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for a in "32 32 128 plain" "64 32 128 plain" "64 64 64 plain"; do timeout 120 python tools/halo_timeline.py $a 2>&1 | grep -v Warn | tail -9; done
echo "--- MAXC=256 NOUT=1"
( B200UNET_HALO_MAXC=256 B200UNET_HALO_NOUT=1 timeout 600 python tools/conv_bench.py all 10 2>&1 | grep convbench ) | grep -E '"ci": (128|256)' | cut -c1-200
echo "--- MAXC=256 NOUT=1 TD=4"
( B200UNET_HALO_MAXC=256 B200UNET_HALO_NOUT=1 B200UNET_HALO_TD=4 timeout 600 python tools/conv_bench.py all 10 2>&1 | grep convbench ) | grep -E '"ci": (128|256)' | cut -c1-200
B200UNET_HALO_MAXC=256 B200UNET_HALO_NOUT=1 timeout 120 python tools/halo_timeline.py 128 128 64 plain 2>&1 | grep -v Warn | tail -8

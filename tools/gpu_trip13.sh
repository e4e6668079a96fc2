#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/gpu_diag.py conv > gpurun_out/diag13.log 2>&1
python - <<PY
import re
bad=[]; n=0
for l in open('gpurun_out/diag13.log'):
    if l.startswith('[diag]'):
        n+=1
        vals=[float(v) for v in re.findall(r'(?:rel|_rel)=([0-9.e+-]+)',l)]
        if 'EXCEPTION' in l or any(v>4e-3 for v in vals): bad.append(l[:260])
print("cases",n,"bad",len(bad)); print("\n".join(bad[:12]))
PY
grep -E "Error|error|timed out" gpurun_out/diag13.log | head -5
for a in "32 32 128 plain" "32 64 128 mode1" "32 32 128 res"; do timeout 120 python tools/halo_timeline.py $a 2>&1 | grep -v Warn | tail -7; done
( timeout 600 python tools/conv_bench.py all 10 2>&1 | grep convbench ) | tee gpurun_out/convbench13.log | cut -c1-330
timeout 300 python tools/layer_times.py gpurun_out/layer_times13.csv 2>&1 | tail -3 | cut -c1-400
( timeout 400 python tools/gpu_diag.py model 2>&1 | grep diag ) | cut -c1-200
echo "--- MAXC=256"
( B200UNET_HALO_MAXC=256 timeout 600 python tools/conv_bench.py all 10 2>&1 | grep convbench ) | grep -E '"ci": (128|256)|total' | cut -c1-330
B200UNET_HALO_MAXC=256 timeout 120 python tools/halo_timeline.py 128 128 64 plain 2>&1 | grep -v Warn | tail -8

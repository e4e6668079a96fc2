#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py conv > gpurun_out/diag_conv6.log 2>&1
python - <<PY
import re
bad=[]; n=0
for l in open('gpurun_out/diag_conv6.log'):
    if l.startswith('[diag] conv'):
        n+=1
        m=re.search(r'(?:rel|dz_rel)=([0-9.e+-]+)',l)
        if 'EXCEPTION' in l or (m and float(m.group(1))>2e-3): bad.append(l[:150])
print("conv cases",n,"bad",len(bad)); print("\n".join(bad[:6]))
PY
( timeout 600 python tools/conv_bench.py all 10 2>&1 | grep convbench ) | tee gpurun_out/convbench6.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo8 python tools/conv_bench.py fwd 1 > gpurun_out/ncu_halo.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 7 -c 1 -o gpurun_out/prof_halo32 python tools/conv_bench.py fwd 1 >> gpurun_out/ncu_halo.log 2>&1
ls -la gpurun_out/*.ncu-rep

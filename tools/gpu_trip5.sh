#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for HS in 1 2 3 6; do
  export B200UNET_HALO_HSPLIT=$HS
  rm -f gpurun_out/diag.log
  timeout 600 python tools/gpu_diag.py conv wgrad > gpurun_out/diag_hs$HS.log 2>&1
  python - <<PY
import re
bad=[]; n=0
for l in open('gpurun_out/diag_hs$HS.log'):
    if l.startswith('[diag] conv') or l.startswith('[diag] wgrad'):
        n+=1
        m=re.search(r'(?:rel|dz_rel)=([0-9.e+-]+)',l)
        if 'EXCEPTION' in l or (m and float(m.group(1))>2e-3): bad.append(l[:150])
print("HSPLIT=$HS cases",n,"bad",len(bad)); print("\n".join(bad[:6]))
PY
  ( timeout 600 python tools/conv_bench.py all 10 2>&1 | grep convbench | grep -E "weighted|\"r\": 128|\"r\": 64" ) | tee gpurun_out/convbench_hs$HS.log
done
unset B200UNET_HALO_HSPLIT
( timeout 300 python tools/gpu_diag.py bench model 2>&1 | grep diag ) | cut -c1-260 | tee gpurun_out/bench_diag.log

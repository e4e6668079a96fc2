#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py elementwise conv > gpurun_out/diag9.log 2>&1
python - <<PY
import re
bad=[]; n=0
for l in open('gpurun_out/diag9.log'):
    if l.startswith('[diag]'):
        n+=1
        vals=[float(v) for v in re.findall(r'(?:rel|_rel)=([0-9.e+-]+)',l)]
        if 'EXCEPTION' in l or any(v>4e-3 for v in vals): bad.append(l[:220])
print("cases",n,"bad",len(bad)); print("\n".join(bad[:8]))
PY
( timeout 600 python tools/conv_bench.py fwd 10 2>&1 | grep convbench ) | tee gpurun_out/convbench9.log | cut -c1-200
timeout 300 python tools/layer_times.py gpurun_out/layer_times9.csv 2>&1 | tail -42 | cut -c1-170
( timeout 300 python tools/gpu_diag.py model 2>&1 | grep diag ) | cut -c1-230

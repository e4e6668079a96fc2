#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py conv > gpurun_out/diag_conv7.log 2>&1
python - <<PY
import re
bad=[]; n=0
for l in open('gpurun_out/diag_conv7.log'):
    if l.startswith('[diag] conv'):
        n+=1
        m=re.search(r'(?:rel|dz_rel)=([0-9.e+-]+)',l)
        m2=re.search(r'(?:stats_rel|bstats_rel)=([0-9.e+-]+)',l)
        if 'EXCEPTION' in l or (m and float(m.group(1))>2e-3) or (m2 and float(m2.group(1))>1e-3): bad.append(l[:200])
print("conv cases",n,"bad",len(bad)); print("\n".join(bad[:8]))
PY
grep -E "EXCEPTION|Error" gpurun_out/diag_conv7.log | head -5
( timeout 600 python tools/conv_bench.py all 10 2>&1 | grep convbench ) | tee gpurun_out/convbench7.log
( timeout 300 python tools/gpu_diag.py bench model 2>&1 | grep diag ) | cut -c1-260 | tee gpurun_out/bench_diag7.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench7.err | tail -1 ) > gpurun_out/bench7.json
python -c "
import json; d=json.load(open('gpurun_out/bench7.json')); print('value',d['value'],'e2e',d['e2e']['value'],'roof',d['roofline']['frac']); print(json.dumps(d['kernels'],indent=0))"

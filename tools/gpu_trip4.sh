#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/diag.log
for g in conv wgrad model; do
  echo "######## $g" >> gpurun_out/diag.log
  timeout 600 python tools/gpu_diag.py $g >> gpurun_out/diag.log 2>&1; echo "exit=$?" >> gpurun_out/diag.log
done
grep -E "EXCEPTION|exit=|wgrad|model" gpurun_out/diag.log | grep -E "diag|exit|EXC" | cut -c1-260 | tail -40
python - <<'PY'
import re
bad=[l for l in open('gpurun_out/diag.log') if l.startswith('[diag] conv') and (('rel=' in l and float(re.search(r'rel=([0-9.e+-]+)',l).group(1))>2e-3))]
print("conv cases with rel>2e-3:", len(bad)); print("".join(bad[:10]))
PY
( timeout 900 python tools/conv_bench.py all 10 2>&1 | grep convbench ) | tee gpurun_out/convbench_all.log
( B200UNET_NO_HALO=1 B200UNET_NO_HALO_WGRAD=1 timeout 900 python tools/conv_bench.py all 10 2>&1 | grep convbench ) | tee gpurun_out/convbench_all_stream.log
( B200UNET_HALO_MAXC=4096 timeout 900 python tools/conv_bench.py fwd 10 2>&1 | grep convbench ) | tee gpurun_out/convbench_halo_all.log
( timeout 300 python tools/gpu_diag.py bench 2>&1 | grep diag ) | tee gpurun_out/bench_diag.log

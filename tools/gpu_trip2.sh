#!/bin/bash
# pytest -m gpu, smoke, bench, cuDNN bar, ncu launch list + full captures of the two tensor-core kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
echo "pytest done: $(tail -1 gpurun_out/pytest_gpu.log)"
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
cat gpurun_out/smoke.log
( timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench.json
cat gpurun_out/bench.json | cut -c1-3000
( timeout 600 python tools/cudnn_baseline.py 5 2>&1 | grep variant ) > gpurun_out/cudnn.jsonl
cat gpurun_out/cudnn.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py 2 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/launches.csv)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_igemm_conv -s 60 -c 12 -o gpurun_out/prof_igemm python tools/prof_step.py 2 > gpurun_out/ncu_igemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wgrad -s 36 -c 6 -o gpurun_out/prof_wgrad python tools/prof_step.py 2 > gpurun_out/ncu_wgrad.log 2>&1
ls -la gpurun_out | tail -15

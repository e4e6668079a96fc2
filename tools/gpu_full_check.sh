#!/bin/bash
# full GPU regression: parity tests, smoke, bf16 gradient study, 1-GPU bench lines for C2 / C3 / C5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 2400 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -60 > gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
timeout 900 python tools/bf16_grad_study.py > gpurun_out/${TAG}_bf16_study.log 2>&1; tail -c 1500 gpurun_out/${TAG}_bf16_study.log
for c in C2 C3 C5; do
  timeout 900 python bench.py --config $c 2>gpurun_out/${TAG}_bench_$c.err > gpurun_out/${TAG}_bench_$c.json
  tail -3 gpurun_out/${TAG}_bench_$c.err; cut -c1-1800 gpurun_out/${TAG}_bench_$c.json
done

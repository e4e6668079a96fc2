#!/bin/bash
# full GPU regression: parity tests, smoke, 1-GPU bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_line.json | cut -c1-1500

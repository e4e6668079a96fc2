#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/conv_determinism.py 3 2>&1 | grep "\[det\]" | grep -E "RACE|bad"
timeout 300 python tools/determinism_check.py 128 2>&1 | grep -v Warn | sed -n 1,4p
( timeout 600 python tools/conv_bench.py all 10 2>&1 | grep convbench ) | tee gpurun_out/convbench13.log | grep total
timeout 300 python tools/layer_times.py gpurun_out/layer_times13.csv 2>&1 | tail -1 | cut -c1-400
timeout 600 python bench.py 2>gpurun_out/bench_err.log | tee gpurun_out/bench_line.json | cut -c1-700

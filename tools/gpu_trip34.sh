#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/conv_determinism.py 3 2>&1 | grep "\[det\]" | grep -E "RACE|bad"
timeout 300 python tools/determinism_check.py 128 2>&1 | grep -v Warn | sed -n 1,6p
bash tools/gpu_trip13.sh 2>&1 | grep -E "bad|steady|convbench\]|wgrad':" | grep -v MAXC | head -22

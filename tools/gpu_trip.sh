#!/bin/bash
# Runs the diagnostic groups each under its own timeout so a hung kernel cannot eat the gpurun budget.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for g in "$@"; do
  echo "######## $g" | tee -a gpurun_out/diag.log
  timeout 420 python tools/gpu_diag.py $g >> gpurun_out/diag.log 2>&1
  echo "exit=$?" | tee -a gpurun_out/diag.log
done
grep -E "^\[diag\]|EXCEPTION|exit=|########|Error|error" gpurun_out/diag.log | tail -150

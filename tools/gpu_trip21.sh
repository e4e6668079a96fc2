#!/bin/bash
cd "$(dirname "$0")/.."
bash tools/gpu_trip19.sh 2>&1 | head -50
bash tools/gpu_trip13.sh

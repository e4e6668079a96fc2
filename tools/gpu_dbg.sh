#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02u}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep "smoke ok\|Error" | tail -3
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -1 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log
grep "dgrad encoder.downsampling" gpurun_out/${TAG}_layer_times.csv

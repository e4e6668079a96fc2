#!/bin/bash
# last GPU call of the round: the parity suite and smoke() on the final sources (deep ring on by default)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02z}
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log

#!/bin/bash
# final regression of the round: parity tests + smoke + the driver's bench line in the default state, then the same tests and the
# bench line with programmatic dependent launch (B200UNET_PDL=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02x}
timeout 1200 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/${TAG}_pytest.log 2>&1; tail -22 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; grep "smoke\[" gpurun_out/${TAG}_smoke.log | cut -c1-260; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py 2>gpurun_out/${TAG}_bench_C2.err > gpurun_out/${TAG}_bench_C2.json; tail -2 gpurun_out/${TAG}_bench_C2.err; cut -c1-400 gpurun_out/${TAG}_bench_C2.json
export B200UNET_PDL=1
timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/${TAG}_bench_C2_pdl.err > gpurun_out/${TAG}_bench_C2_pdl.json; tail -2 gpurun_out/${TAG}_bench_C2_pdl.err; cut -c1-400 gpurun_out/${TAG}_bench_C2_pdl.json
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_pdl.log 2>&1; tail -8 gpurun_out/${TAG}_pytest_pdl.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke_pdl.log 2>&1; tail -1 gpurun_out/${TAG}_smoke_pdl.log

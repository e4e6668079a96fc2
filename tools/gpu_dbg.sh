#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02g}
timeout 900 python tools/graph_debug.py > gpurun_out/${TAG}_graph_debug.log 2>&1; cat gpurun_out/${TAG}_graph_debug.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
for c in C2 C3 C5; do
  timeout 900 python bench.py --config $c > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err; tail -c 600 gpurun_out/${TAG}_bench_$c.err; cut -c1-700 gpurun_out/${TAG}_bench_$c.json
done
timeout 600 python bench.py --config C2 --no-graph > gpurun_out/${TAG}_bench_C2_eager.json 2> gpurun_out/${TAG}_bench_C2_eager.err; cut -c1-400 gpurun_out/${TAG}_bench_C2_eager.json
timeout 600 python tools/dynunet_bench.py > gpurun_out/${TAG}_dynunet_bench.json 2> gpurun_out/${TAG}_dynunet_bench.err; cat gpurun_out/${TAG}_dynunet_bench.json

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02t}
for c in C2 C3 C5; do
  timeout 900 python bench.py --config $c > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err; tail -c 300 gpurun_out/${TAG}_bench_$c.err; cut -c1-330 gpurun_out/${TAG}_bench_$c.json
done
timeout 600 python bench.py --config C2 --no-graph --no-cpu-baseline > gpurun_out/${TAG}_bench_C2_eager.json 2> gpurun_out/${TAG}_bench_C2_eager.err; cut -c1-330 gpurun_out/${TAG}_bench_C2_eager.json
timeout 600 python tools/dynunet_bench.py > gpurun_out/${TAG}_dynunet_bench.json 2> gpurun_out/${TAG}_dynunet_bench.err; cat gpurun_out/${TAG}_dynunet_bench.json
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -1 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log

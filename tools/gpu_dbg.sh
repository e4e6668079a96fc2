#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02e}
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log; tail -8 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/conv_bench.py epi 10 > gpurun_out/${TAG}_convbench_lean.log 2>&1; grep weighted gpurun_out/${TAG}_convbench_lean.log
B200UNET_HALO_GENERIC_EPILOGUE=1 timeout 600 python tools/conv_bench.py epi 10 > gpurun_out/${TAG}_convbench_generic.log 2>&1; grep weighted gpurun_out/${TAG}_convbench_generic.log
timeout 600 python tools/layer_times.py gpurun_out/${TAG}_layer_times.csv > gpurun_out/${TAG}_layer_times.log 2>&1; head -1 gpurun_out/${TAG}_layer_times.log; tail -1 gpurun_out/${TAG}_layer_times.log
for c in C2 C3 C5; do
  timeout 900 python bench.py --config $c 2>gpurun_out/${TAG}_bench_$c.err > gpurun_out/${TAG}_bench_$c.json
  tail -2 gpurun_out/${TAG}_bench_$c.err; cut -c1-400 gpurun_out/${TAG}_bench_$c.json
done
timeout 600 python tools/dynunet_bench.py > gpurun_out/${TAG}_dynunet.json 2>gpurun_out/${TAG}_dynunet.err; cat gpurun_out/${TAG}_dynunet.json; tail -2 gpurun_out/${TAG}_dynunet.err

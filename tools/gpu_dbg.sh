cd /root/repo 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_prepost.py -x -q -k "graphed_train_step" 2>&1 | tail -120 > gpurun_out/r02b_graph.log
timeout 600 python -m pytest tests/test_gpu_dynunet.py -x -q -k "1-2-filters1" 2>&1 | tail -80 > gpurun_out/r02b_dyn.log
timeout 600 python -m pytest tests/test_gpu_prepost.py -q -k "second_forward or forward_only" 2>&1 | tail -80 > gpurun_out/r02b_misc.log
timeout 900 python bench.py --config C5 2>gpurun_out/r02b_bench_C5.err > gpurun_out/r02b_bench_C5.json; tail -5 gpurun_out/r02b_bench_C5.err; cut -c1-600 gpurun_out/r02b_bench_C5.json
timeout 900 python bench.py --config C2 --no-graph 2>gpurun_out/r02b_bench_C2_nograph.err > gpurun_out/r02b_bench_C2_nograph.json; tail -5 gpurun_out/r02b_bench_C2_nograph.err; cut -c1-900 gpurun_out/r02b_bench_C2_nograph.json
tools/gpu_experiments.sh r02b

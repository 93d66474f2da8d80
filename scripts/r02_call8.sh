set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_leaf.py tests/test_gpu_vs_reference_kernels.py tests/test_gpu_baseline_size.py -m gpu -q -x -k "chamfer" > gpurun_out/r02_c8_tests.log 2>&1
tail -5 gpurun_out/r02_c8_tests.log
timeout 300 python scripts/bench_nn_ab.py --big > gpurun_out/r02_c8_nn_ab.jsonl 2> gpurun_out/r02_c8_nn.err
cat gpurun_out/r02_c8_nn_ab.jsonl
tail -3 gpurun_out/r02_c8_nn.err

set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_deflowpp.py -m gpu -q -x > gpurun_out/r02_c6_tests.log 2>&1
tail -3 gpurun_out/r02_c6_tests.log
timeout 100 python scripts/trace_conv_tiles.py 24 > gpurun_out/r02_c6_trace_enc2_x24.txt 2>&1
sed -n 5,8p gpurun_out/r02_c6_trace_enc2_x24.txt; tail -1 gpurun_out/r02_c6_trace_enc2_x24.txt
timeout 100 python scripts/trace_conv_tiles.py 12 256 64 1 > gpurun_out/r02_c6_trace_enc1_x12.txt 2>&1
sed -n 5,8p gpurun_out/r02_c6_trace_enc1_x12.txt; tail -1 gpurun_out/r02_c6_trace_enc1_x12.txt
timeout 200 python scripts/bench_conv.py 2 > gpurun_out/r02_c6_conv_layers.txt 2>&1
cat gpurun_out/r02_c6_conv_layers.txt
timeout 200 python bench.py --warmup 3 > gpurun_out/r02_c6_bench.json 2> gpurun_out/r02_c6_bench.err
cat gpurun_out/r02_c6_bench.json

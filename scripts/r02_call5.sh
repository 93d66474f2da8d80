set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q > gpurun_out/r02_c5_tests.log 2>&1
tail -3 gpurun_out/r02_c5_tests.log
timeout 100 python scripts/trace_conv_tiles.py 3 > gpurun_out/r02_c5_trace_enc2_x3.txt 2>&1
timeout 100 python scripts/trace_conv_tiles.py 24 > gpurun_out/r02_c5_trace_enc2_x24.txt 2>&1
timeout 100 python scripts/trace_conv_tiles.py 12 256 64 1 > gpurun_out/r02_c5_trace_enc1_x12.txt 2>&1
head -30 gpurun_out/r02_c5_trace_enc2_x24.txt
timeout 200 python bench.py --warmup 3 --no-cpu-baseline > gpurun_out/r02_c5_bench.json 2> gpurun_out/r02_c5_bench.err
cat gpurun_out/r02_c5_bench.json

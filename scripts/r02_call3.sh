set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_deflowpp.py tests/test_gpu_conv.py tests/test_gpu_baseline_size.py tests/test_gpu_fastnsf.py tests/test_gpu_engine.py -m gpu -q -x > gpurun_out/r02_c3_tests.log 2>&1
tail -5 gpurun_out/r02_c3_tests.log
timeout 200 python bench.py --warmup 3 > gpurun_out/r02_c3_bench_default.json 2> gpurun_out/r02_c3_bench.err
cat gpurun_out/r02_c3_bench_default.json
HIMO_PDL=0 timeout 200 python bench.py --warmup 3 --no-cpu-baseline > gpurun_out/r02_c3_bench_nopdl.json 2>> gpurun_out/r02_c3_bench.err
cat gpurun_out/r02_c3_bench_nopdl.json
HIMO_COMPOSE_SKIP=0 timeout 200 python bench.py --warmup 3 --no-cpu-baseline > gpurun_out/r02_c3_bench_nocompose.json 2>> gpurun_out/r02_c3_bench.err
cat gpurun_out/r02_c3_bench_nocompose.json
tail -5 gpurun_out/r02_c3_bench.err

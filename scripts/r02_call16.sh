set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fastnsf.py tests/test_gpu_nsfp.py tests/test_gpu_baseline_size.py -m gpu -q -x -k "not chamfer and not seflowpp" > gpurun_out/r02_c16_tests.log 2>&1
tail -12 gpurun_out/r02_c16_tests.log
timeout 300 python scripts/bench_fastnsf.py > gpurun_out/r02_c16_fastnsf.json 2> gpurun_out/r02_c16_fastnsf.err
cat gpurun_out/r02_c16_fastnsf.json; tail -3 gpurun_out/r02_c16_fastnsf.err
timeout 200 python scripts/bench_nsfp.py > gpurun_out/r02_c16_nsfp.json 2>> gpurun_out/r02_c16_fastnsf.err
cat gpurun_out/r02_c16_nsfp.json

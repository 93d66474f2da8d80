set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_nsfp.py tests/test_gpu_fastnsf.py tests/test_gpu_cli.py -m gpu -q -x > gpurun_out/r02_c9_tests.log 2>&1
tail -15 gpurun_out/r02_c9_tests.log
timeout 200 python scripts/bench_nsfp.py > gpurun_out/r02_c9_nsfp_bench.json 2> gpurun_out/r02_c9_nsfp.err
cat gpurun_out/r02_c9_nsfp_bench.json; tail -3 gpurun_out/r02_c9_nsfp.err

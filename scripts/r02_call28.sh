set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fastnsf.py tests/test_gpu_baseline_size.py -m gpu -q -x 2>&1 | tail -8
timeout 200 python scripts/bench_fastnsf.py > gpurun_out/r02_c28_fastnsf.json 2> gpurun_out/r02_c28_fastnsf.err; cat gpurun_out/r02_c28_fastnsf.json; tail -3 gpurun_out/r02_c28_fastnsf.err

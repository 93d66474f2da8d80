set -u
mkdir -p gpurun_out
timeout 300 python scripts/bench_dt.py 2>&1 | tail -1 | tee gpurun_out/r02_c30_dt.json
timeout 900 python -m pytest tests/test_gpu_fastnsf.py tests/test_gpu_baseline_size.py tests/test_gpu_cli.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python scripts/bench_fastnsf.py > gpurun_out/r02_c30_fastnsf.json 2> /dev/null; cat gpurun_out/r02_c30_fastnsf.json

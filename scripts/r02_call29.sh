set -u
mkdir -p gpurun_out
timeout 300 python scripts/bench_dt.py 2>&1 | tail -2 | tee gpurun_out/r02_c29_dt.json
timeout 600 python -m pytest tests/test_gpu_fastnsf.py -m gpu -q -x 2>&1 | tail -4

set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fastnsf.py tests/test_gpu_nsfp.py tests/test_gpu_baseline_size.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python scripts/bench_fastnsf.py 2>/dev/null | cut -c1-330
timeout 200 python scripts/bench_nsfp.py 2>/dev/null | tail -1

set -u
mkdir -p gpurun_out
timeout 60 ./scripts/exp/mn_major > gpurun_out/r02_c10_mn_major.txt 2>&1
cat gpurun_out/r02_c10_mn_major.txt
timeout 400 python -m pytest tests/test_gpu_cli.py -m gpu -q -x > gpurun_out/r02_c10_tests.log 2>&1
tail -5 gpurun_out/r02_c10_tests.log

set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_baseline_size.py tests/test_dt_edt_bound.py tests/test_gpu_fastnsf.py tests/test_gpu_engine.py tests/test_gpu_cli.py -m gpu -q -x > gpurun_out/r02_newtests.log 2>&1
tail -5 gpurun_out/r02_newtests.log
timeout 200 python scripts/bench_conv.py 2 > gpurun_out/r02_conv_probe.txt 2>&1
cat gpurun_out/r02_conv_probe.txt

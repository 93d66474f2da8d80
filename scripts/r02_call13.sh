set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/r02_c13_tests.log 2>&1
tail -3 gpurun_out/r02_c13_tests.log
TSTORE=0 timeout 100 python scripts/bench_conv.py 2 "mlp" > gpurun_out/r02_c13_ts.txt 2>&1
TSTORE=1 timeout 100 python scripts/bench_conv.py 2 "mlp" >> gpurun_out/r02_c13_ts.txt 2>&1
TSTORE=0 timeout 100 python scripts/bench_conv.py 2 "enc2.0" >> gpurun_out/r02_c13_ts.txt 2>&1
cat gpurun_out/r02_c13_ts.txt
timeout 300 python scripts/bench_fastnsf.py > gpurun_out/r02_c13_fastnsf.json 2> gpurun_out/r02_c13_fastnsf.err
cat gpurun_out/r02_c13_fastnsf.json

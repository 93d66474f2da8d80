set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/r02_c12_tests.log 2>&1
tail -5 gpurun_out/r02_c12_tests.log
TSTORE=0 timeout 100 python scripts/bench_conv.py 2 "enc2" > gpurun_out/r02_c12_ts_off.txt 2>&1
TSTORE=0 timeout 100 python scripts/bench_conv.py 2 "mlp" >> gpurun_out/r02_c12_ts_off.txt 2>&1
TSTORE=1 timeout 100 python scripts/bench_conv.py 2 "enc2" > gpurun_out/r02_c12_ts_on.txt 2>&1
TSTORE=1 timeout 100 python scripts/bench_conv.py 2 "mlp" >> gpurun_out/r02_c12_ts_on.txt 2>&1
paste -d'\n' gpurun_out/r02_c12_ts_off.txt gpurun_out/r02_c12_ts_on.txt
timeout 600 python -m pytest tests/test_gpu_deflowpp.py tests/test_gpu_fastnsf.py tests/test_gpu_nsfp.py -m gpu -q -x > gpurun_out/r02_c12_tests2.log 2>&1
tail -5 gpurun_out/r02_c12_tests2.log
timeout 300 python scripts/bench_fastnsf.py > gpurun_out/r02_c12_fastnsf.json 2> gpurun_out/r02_c12_fastnsf.err
cat gpurun_out/r02_c12_fastnsf.json

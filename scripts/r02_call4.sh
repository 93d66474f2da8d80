set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x > gpurun_out/r02_c4_tests.log 2>&1
tail -5 gpurun_out/r02_c4_tests.log
ROWS2=0 timeout 100 python scripts/bench_conv.py 2 "b" > gpurun_out/r02_c4_rows2_off.txt 2>&1
ROWS2=1 timeout 100 python scripts/bench_conv.py 2 "b" > gpurun_out/r02_c4_rows2_on.txt 2>&1
ROWS2=1 timeout 100 python scripts/bench_conv.py 2 "dec4" >> gpurun_out/r02_c4_rows2_on.txt 2>&1
ROWS2=0 timeout 100 python scripts/bench_conv.py 2 "dec4" >> gpurun_out/r02_c4_rows2_off.txt 2>&1
paste -d'\n' gpurun_out/r02_c4_rows2_off.txt gpurun_out/r02_c4_rows2_on.txt
timeout 300 python -m pytest tests/test_gpu_deflowpp.py tests/test_gpu_baseline_size.py -m gpu -q -x -k "not chamfer and not fastnsf" > gpurun_out/r02_c4_tests2.log 2>&1
tail -5 gpurun_out/r02_c4_tests2.log
timeout 200 python bench.py --warmup 3 > gpurun_out/r02_c4_bench.json 2> gpurun_out/r02_c4_bench.err
cat gpurun_out/r02_c4_bench.json

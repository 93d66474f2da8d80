set -u
mkdir -p gpurun_out
( time timeout 600 python bench.py --warmup 3 > gpurun_out/r02_c7_bench.json 2> gpurun_out/r02_c7_bench.err ) 2> gpurun_out/r02_c7_time.txt
cat gpurun_out/r02_c7_bench.json
tail -5 gpurun_out/r02_c7_bench.err
cat gpurun_out/r02_c7_time.txt
timeout 120 python scripts/bench_ref_kernels.py > gpurun_out/r02_c7_ref_kernels.json 2>> gpurun_out/r02_c7_bench.err
cat gpurun_out/r02_c7_ref_kernels.json

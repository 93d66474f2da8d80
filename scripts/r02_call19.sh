set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_c19_tests.log 2>&1
tail -15 gpurun_out/r02_c19_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c19_smoke.log 2>&1; echo smoke=$?
tail -3 gpurun_out/r02_c19_smoke.log
timeout 600 python bench.py > gpurun_out/r02_c19_bench.json 2> gpurun_out/r02_c19_bench.err; echo bench=$?
cat gpurun_out/r02_c19_bench.json | head -c 6000; tail -3 gpurun_out/r02_c19_bench.err

#!/usr/bin/env bash
# One GPU call that re-establishes the measured state of the repo (about 4 minutes on a B200):
#   gpurun --timeout 420 -- 'bash scripts/gpu_checkin.sh'
# 1. the opt-in comparison with the reference's own Chamfer kernel (strict equality after the rounding-sequence fix),
# 2. the whole -m gpu suite, 3. smoke(), 4. the bench line, 5. the side-by-side bench against the reference kernels.
set -u
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_vs_reference_kernels.py -m gpu -q > gpurun_out/checkin_ref_kernels.log 2>&1
tail -3 gpurun_out/checkin_ref_kernels.log
timeout 240 python -m pytest tests -m gpu -q > gpurun_out/checkin_gpu_tests.log 2>&1
tail -3 gpurun_out/checkin_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/checkin_smoke.log 2>&1
tail -1 gpurun_out/checkin_smoke.log
timeout 180 python bench.py --warmup 3 > gpurun_out/checkin_bench.json 2> gpurun_out/checkin_bench.err
cat gpurun_out/checkin_bench.json
timeout 60 python scripts/bench_ref_kernels.py > gpurun_out/checkin_ref_kernels.json 2>> gpurun_out/checkin_bench.err
cat gpurun_out/checkin_ref_kernels.json

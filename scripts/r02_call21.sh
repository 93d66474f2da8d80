set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_cli.py tests/test_gpu_fastnsf.py -m gpu -q -x > gpurun_out/r02_c21_tests.log 2>&1
tail -5 gpurun_out/r02_c21_tests.log
timeout 300 python scripts/exp_two_streams.py > gpurun_out/r02_c21_streams.json 2> gpurun_out/r02_c21_streams.err
cat gpurun_out/r02_c21_streams.json; tail -3 gpurun_out/r02_c21_streams.err
timeout 600 python bench.py > gpurun_out/r02_c21_bench.json 2> gpurun_out/r02_c21_bench.err; echo bench=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c21_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','in_flight','stages_ms','sustained','pipeline','clocks'):
    print(k, d.get(k))
print(d['fastnsf'].get('engine'), d['fastnsf'].get('configured_run'))
PY
tail -3 gpurun_out/r02_c21_bench.err

set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fastnsf.py tests/test_gpu_baseline_size.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02_c38_bench.json 2> gpurun_out/r02_c38_bench.err; echo bench=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c38_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','stages_ms','clocks'):
    print(k, d.get(k))
print('fastnsf', {k: d['fastnsf'].get(k) for k in ('ms_per_iter','dt_build_ms','configured_run','engine','engine_stream')})
print('knn', json.dumps(d['knn'])[:1500])
PY
tail -3 gpurun_out/r02_c38_bench.err

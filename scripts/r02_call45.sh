set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_c45_tests.log 2>&1
tail -4 gpurun_out/r02_c45_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_c45_bench.json 2> gpurun_out/r02_c45_bench.err; echo bench=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c45_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','stages_ms','clocks'):
    print(k, d.get(k))
print('single', d['in_flight']['single_stream'])
print('roofline', {k: d['roofline'].get(k) for k in ('achieved','frac','traffic')})
print('sustained', d['sustained'].get('frames_per_s'), 'pipeline', d['pipeline'].get('save_frames_per_s'))
print('fastnsf', {k: d['fastnsf'].get(k) for k in ('ms_per_iter','dt_build_ms','configured_run','engine','engine_stream','algorithmic_tflops')})
PY
tail -3 gpurun_out/r02_c45_bench.err

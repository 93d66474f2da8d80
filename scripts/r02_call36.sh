set -u
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_c36_bench.json 2> gpurun_out/r02_c36_bench.err; echo bench=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c36_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','in_flight','stages_ms','clocks','cpu_baseline','epe_vs_cpu_reference'):
    print(k, d.get(k))
print('roofline', {k: d['roofline'].get(k) for k in ('achieved','frac','traffic','time_basis','single_stream','frac_of_sustained_peak')})
print('sustained', d.get('sustained')); print('pipeline', d.get('pipeline'))
print('fastnsf', {k: d['fastnsf'].get(k) for k in ('ms_per_iter','dt_build_ms','configured_run','engine','engine_stream','algorithmic_tflops')})
print('knn', d['knn'].get('lidar_100k'), d['knn'].get('uniform_1m')); print('voxelize', d['voxelize'])
PY
tail -3 gpurun_out/r02_c36_bench.err

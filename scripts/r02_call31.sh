set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus ${NG:-2} --steps 50 --warmup 3 > gpurun_out/r02_c31_bench2.json 2> gpurun_out/r02_c31_bench2.err; echo bench2=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c31_bench2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','n_gpus','e2e','in_flight','stages_ms','clocks'):
    print(k, d.get(k))
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','time_basis','single_stream')})
print('sustained', d.get('sustained')); print('pipeline', d.get('pipeline'))
print('fastnsf', {k: d['fastnsf'].get(k) for k in ('ms_per_iter','dt_build_ms','configured_run','engine')})
print('knn', d['knn'].get('lidar_100k'), d['knn'].get('uniform_1m'))
PY
tail -3 gpurun_out/r02_c31_bench2.err


set -u
mkdir -p gpurun_out
CASES='[[4,0,1,100000,150,0,"",{}],[4,1,1,100000,150,0,"",{}],[3,1,1,100000,150,0,"",{}],[4,1,1,100000,150,0,"",{}],[2,1,1,100000,100,0,"",{}],[1,1,1,100000,100,0,"",{}],[3,1,1,100000,150,0,"",{}],[6,1,1,100000,150,0,"",{}]]'
timeout 700 python scripts/exp_lanes_bisect.py "$CASES" 2>&1 | tee gpurun_out/r02_c26_lanes.txt
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_deflowpp.py tests/test_gpu_engine.py tests/test_gpu_fastnsf.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02_c26_bench.json 2> gpurun_out/r02_c26_bench.err; echo bench=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c26_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e','in_flight','stages_ms','sustained','pipeline','clocks'):
    print(k, d.get(k))
print(d['fastnsf'].get('engine'), d['fastnsf'].get('configured_run'), d['fastnsf'].get('ms_per_iter'))
PY
tail -3 gpurun_out/r02_c26_bench.err

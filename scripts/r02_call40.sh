set -u
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout 120 python bench.py --no-extras --no-cpu-baseline --steps 300 > gpurun_out/r02_c40_tmp.json 2>/dev/null; rc=$?
  python - "$i" "$rc" <<'PY' | tee -a gpurun_out/r02_c40_stress.txt
import json, sys
try:
    d=json.loads(open('gpurun_out/r02_c40_tmp.json').read().strip().splitlines()[-1])
    print("run", sys.argv[1], "rc", sys.argv[2], "value %.1f e2e %.1f single %.1f clocks %s" % (d['value'], d['e2e']['value'], d['in_flight']['single_stream']['value'], d['clocks'].get('sm_mhz')))
except Exception as e:
    print("run", sys.argv[1], "rc", sys.argv[2], "FAILED", e)
PY
done
HIMO_NSF_WORKERS=3 timeout 300 python - <<'PY' 2>&1 | tail -3
import time, torch, bench
from himo_b200.engine import FastNSFEngine
fr = bench.make_frames(0, 1)[0]
eng = FastNSFEngine(device="cuda:0", itr_num=200, early_patience=10)
t0 = time.perf_counter()
n = 0
for _ in eng.infer_stream(fr for _ in range(40)):
    n += 1
print("fastnsf stream pairs", n, "ms/pair", (time.perf_counter() - t0) / n * 1e3)
PY

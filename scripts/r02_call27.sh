set -u
mkdir -p gpurun_out
for cfg in "1 3" "0 3" "1 2" "0 2" "1 4" "0 4" "1 3" "0 3"; do
  set -- $cfg
  HIMO_PDL=$1 HIMO_SLOTS=$2 timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 100 > gpurun_out/r02_c27_tmp.json 2>/dev/null
  python - "$1" "$2" <<'PY' | tee -a gpurun_out/r02_c27_ab.txt
import json, sys
d=json.loads(open('gpurun_out/r02_c27_tmp.json').read().strip().splitlines()[-1])
print("pdl", sys.argv[1], "slots", sys.argv[2], "value %.1f e2e %.1f single %.1f backbone_ms %.3f clocks %s" % (d['value'], d['e2e']['value'], d['in_flight']['single_stream']['value'], d['stages_ms']['backbone'], d['clocks'].get('sm_mhz')))
PY
done

set -u
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c47_nsfp.csv -s 600 -c 70 python scripts/bench_nsfp.py > gpurun_out/r02_c47.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.DictReader([ln for ln in open('gpurun_out/r02_c47_nsfp.csv') if ln.startswith('"')])]
agg=collections.OrderedDict()
for r in rows:
    n=r["Kernel Name"].split("(")[0].replace("himo::","").replace("void ","")[:60]
    v=float(r["Metric Value"].replace(",",""))/1e3
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for n,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]:
    print("%6.1f us total %3d launches  %5.1f us each  %s"%(a[1],a[0],a[1]/a[0],n))
print("total", tot, "over", len(rows), "launches")
PY

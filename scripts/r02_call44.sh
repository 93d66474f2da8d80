set -u
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_nsf_launches_fp32.csv python scripts/prof_fastnsf.py fp32 > /dev/null 2>&1
python - <<'PY'
import csv
rows={}
lines=[ln for ln in open('gpurun_out/r02_nsf_launches_fp32.csv') if ln.startswith('"')]
for r in csv.DictReader(lines):
    e=rows.setdefault(int(r["ID"]),{"name":r["Kernel Name"].split("(")[0].replace("himo::","")})
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
    if r["Metric Name"].startswith("gpu__time"): e["us"]=v/1e3 if u.startswith("n") else v
    elif "read" in r["Metric Name"]: e["rd"]=v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[u]
    else: e["wr"]=v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[u]
for k in sorted(rows)[:9]:
    e=rows[k]; print("  %8.1f us rd %7.1f wr %7.1f MB %s"%(e["us"],e["rd"]/1e6,e["wr"]/1e6,e["name"]))
PY

set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for P in fp32 bf16; do
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_nsf_launches_$P.csv python scripts/prof_fastnsf.py $P > gpurun_out/r02_c35_$P.log 2>&1; tail -1 gpurun_out/r02_c35_$P.log
done
python - <<'PY'
import csv
for P in ("fp32","bf16"):
    rows={}
    lines=[ln for ln in open(f'gpurun_out/r02_nsf_launches_{P}.csv') if ln.startswith('"')]
    for r in csv.DictReader(lines):
        e=rows.setdefault(int(r["ID"]),{"name":r["Kernel Name"].split("(")[0].replace("himo::","")})
        v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]
        if r["Metric Name"].startswith("gpu__time"): e["us"]=v/1e3 if u.startswith("n") else v
        elif "read" in r["Metric Name"]: e["rd"]=v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[u]
        else: e["wr"]=v*{"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}[u]
    print(P)
    for k in sorted(rows)[:9]:
        e=rows[k]; print("  %8.1f us rd %7.1f wr %7.1f MB %s"%(e["us"],e["rd"]/1e6,e["wr"]/1e6,e["name"]))
PY

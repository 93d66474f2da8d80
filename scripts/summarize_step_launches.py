"""ncu launch list (csv of gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) of one SeFlow++ step ->
profiles/r02_step_launches_summary.txt and profiles/r02_backbone_traffic.json (stamped with the SHA-256 of csrc/conv.cu)."""
import csv, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_step_launches.csv")
rows = {}
with open(src) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    k = int(r["ID"])
    e = rows.setdefault(k, {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        e["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    elif r["Metric Name"] == "dram__bytes_read.sum":
        e["rd"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    elif r["Metric Name"] == "dram__bytes_write.sum":
        e["wr"] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
out = ["One SeFlow++ step (100k-pt lidar triple), round 2 end state (one network alone on its stream): ncu --profile-from-start off "
       "--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python scripts/prof_step.py",
       "(per-launch times are cold-cache and serialised: compare shares, not absolutes)", ""]
tot = conv = crd = cwr = 0.0
for k in sorted(rows):
    e = rows[k]
    name = e["name"].split("(")[0].replace("himo::", "")
    out.append("%4d  %8.1f us  dram rd %8.1f MB wr %8.1f MB  %s" % (k, e["us"], e["rd"] / 1e6, e["wr"] / 1e6, name))
    tot += e["us"]
    if "k_conv" in name:
        conv += e["us"]; crd += e["rd"]; cwr += e["wr"]
out += ["", "total %.1f us over %d launches; convolution kernels: %.1f us = %.1f %% of the step; their DRAM traffic %.1f MB read + %.1f MB "
        "written per step" % (tot, len(rows), conv, 100 * conv / tot, crd / 1e6, cwr / 1e6)]
open(os.path.join(ROOT, "profiles", "r02_step_launches_summary.txt"), "w").write("\n".join(out) + "\n")
sha = hashlib.sha256(open(os.path.join(ROOT, "himo_b200", "csrc", "conv.cu"), "rb").read()).hexdigest()[:16]
json.dump({"dram_bytes_read_per_step": crd, "dram_bytes_write_per_step": cwr, "conv_cu_sha16": sha,
           "source": "ncu launch list, profiles/r02_step_launches_summary.txt (dram__bytes_read.sum / dram__bytes_write.sum over the "
                     "k_conv_* launches of one step)"},
          open(os.path.join(ROOT, "profiles", "r02_backbone_traffic.json"), "w"), indent=1)
print(out[-1])

set -u
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"k_mlp_chain|k_nsf_dw|k_nsf_head|k_nsf_adam" -c 5 -o /tmp/nsf_full python scripts/prof_fastnsf.py fp32 > gpurun_out/r02_c41.log 2>&1; tail -1 gpurun_out/r02_c41.log
ncu -i /tmp/nsf_full.ncu-rep --page raw --csv > gpurun_out/r02_nsf_full_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_nsf_full_raw.csv')))
hdr=rows[0]; units=rows[1]
want=["Kernel Name","gpu__time_duration.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","dram__bytes_read.sum","dram__bytes_write.sum",
      "sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__inst_executed_pipe_tensor.sum","sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.sum",
      "sm__cycles_elapsed.max","lts__t_sector_hit_rate.pct","l1tex__data_bank_conflicts_pipe_lsu.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active"]
idx=[hdr.index(w) if w in hdr else -1 for w in want]
print(" | ".join(want))
print(" | ".join(units[i] if i>=0 else "-" for i in idx))
for r in rows[2:]:
    print(" | ".join((r[i][:40] if i>=0 else "-") for i in idx))
PY

set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_step_launches.csv python scripts/prof_step.py > gpurun_out/r02_c34_a.log 2>&1; tail -1 gpurun_out/r02_c34_a.log
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_dt_launches.csv python scripts/prof_dt.py > gpurun_out/r02_c34_b.log 2>&1; tail -1 gpurun_out/r02_c34_b.log
wc -l gpurun_out/r02_step_launches.csv gpurun_out/r02_dt_launches.csv

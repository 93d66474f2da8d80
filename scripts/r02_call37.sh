set -u
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_nsf_dt_sweep -c 1 -o /tmp/dt_sweep python scripts/prof_dt.py > gpurun_out/r02_c37.log 2>&1; tail -2 gpurun_out/r02_c37.log
ncu -i /tmp/dt_sweep.ncu-rep --page raw --csv > gpurun_out/r02_dt_sweep_raw.csv 2>/dev/null
ncu -i /tmp/dt_sweep.ncu-rep --page details > gpurun_out/r02_dt_sweep_details.txt 2>/dev/null
ncu -i /tmp/dt_sweep.ncu-rep --page source --csv > gpurun_out/r02_dt_sweep_source.csv 2>/dev/null
ls -la gpurun_out/r02_dt_sweep_*

set -u
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_step_launches.csv python scripts/prof_step.py > gpurun_out/r02_c15_ncu1.log 2>&1
tail -2 gpurun_out/r02_c15_ncu1.log
wc -l gpurun_out/r02_step_launches.csv
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:k_conv -o /tmp/r02_conv_full python scripts/prof_step.py > gpurun_out/r02_c15_ncu2.log 2>&1
tail -2 gpurun_out/r02_c15_ncu2.log
ncu -i /tmp/r02_conv_full.ncu-rep --page raw --csv > gpurun_out/r02_conv_full_raw.csv 2> gpurun_out/r02_c15_ncu3.log
ls -la gpurun_out/r02_conv_full_raw.csv /tmp/r02_conv_full.ncu-rep

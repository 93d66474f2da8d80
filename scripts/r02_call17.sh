set -u
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_fastnsf_launches.csv python scripts/prof_fastnsf.py > gpurun_out/r02_c17_ncu.log 2>&1
tail -2 gpurun_out/r02_c17_ncu.log; wc -l gpurun_out/r02_fastnsf_launches.csv

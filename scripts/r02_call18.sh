set -u
mkdir -p gpurun_out
HIMO_DBG_CHAIN_SKIP_STORE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c18_nostore.csv python scripts/prof_fastnsf.py > gpurun_out/r02_c18.log 2>&1
grep -c "k_mlp_chain" gpurun_out/r02_c18_nostore.csv
grep "k_mlp_chain" gpurun_out/r02_c18_nostore.csv | head -4 | cut -c1-60,200-400

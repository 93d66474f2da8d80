set -u
mkdir -p gpurun_out
for v in 0 1 2; do
HIMO_DBG_CHAIN_STORE=$v timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_c42_$v.csv -k regex:k_mlp_chain python scripts/prof_fastnsf.py fp32 > /dev/null 2>&1
echo "dbg_store=$v"; grep -o 'k_mlp_chain<[01]>.*' gpurun_out/r02_c42_$v.csv | sed 's/(ChainMaps.*Command line profiler metrics//' | head -8
done

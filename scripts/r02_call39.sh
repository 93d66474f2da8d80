set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_deflowpp.py tests/test_gpu_engine.py tests/test_gpu_baseline_size.py -m gpu -q -x 2>&1 | tail -3
M="gpu__time_duration.sum"
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_c39_launches.csv -k regex:"k_dec|k_embed|k_clear" python scripts/prof_step.py > gpurun_out/r02_c39.log 2>&1
grep -o '"k_[a-z_0-9]*[^"]*","1","[0-9]*","([0-9, ]*)","([0-9, ]*)","0","10.0","[^"]*","gpu__time_duration.sum","[a-z]*","[0-9.,]*"' gpurun_out/r02_c39_launches.csv | sed 's/","1",.*duration.sum"//' | head -12

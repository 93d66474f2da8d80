set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_deflowpp.py tests/test_gpu_baseline_size.py -m gpu -q -x 2>&1 | tail -3
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_c46_launches.csv -k regex:"k_upsample|k_dec_gather" python scripts/prof_step.py > gpurun_out/r02_c46.log 2>&1
grep -o '"k_[a-z_0-9]*[^"]*","1","[0-9]*","([0-9, ]*)","([0-9, ]*)","0","10.0","[^"]*","gpu__time_duration.sum","[a-z]*","[0-9.,]*"' gpurun_out/r02_c46_launches.csv | sed 's/","1",.*duration.sum"//' | cut -c1-60,150-

set -u
mkdir -p gpurun_out
timeout 300 python scripts/exp_fastnsf_engine.py > gpurun_out/r02_c20_nsf_engine.json 2> gpurun_out/r02_c20_nsf_engine.err
cat gpurun_out/r02_c20_nsf_engine.json; tail -3 gpurun_out/r02_c20_nsf_engine.err
timeout 300 python scripts/exp_two_streams.py > gpurun_out/r02_c20_two_streams.json 2> gpurun_out/r02_c20_two_streams.err
cat gpurun_out/r02_c20_two_streams.json; tail -3 gpurun_out/r02_c20_two_streams.err

set -x
python -m pytest tests/test_gpu_wide.py tests/test_gpu_fused.py -m gpu -q -x -k "wide or burgers" > gpurun_out/r2u_wide_tests.log 2>&1; tail -3 gpurun_out/r2u_wide_tests.log | cut -c1-300
python scripts/wide_timing.py 4000000 3 > gpurun_out/r2u_wide_timing.log 2>&1; cat gpurun_out/r2u_wide_timing.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2u_wide_launches.csv python scripts/wide_timing.py 1000000 2 > gpurun_out/r2u_ncu.log 2>&1

set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2x_gpu_tests.log 2>&1; tail -3 gpurun_out/r2x_gpu_tests.log
python scripts/wide_timing.py 4000000 4 > gpurun_out/r2x_wide_timing_pf.log 2>&1
MPINN_B200_LIB=$PWD/multipinn_b200/csrc/_variants/libmpinn_b200_base.so python scripts/wide_timing.py 4000000 4 > gpurun_out/r2x_wide_timing_base.log 2>&1
cat gpurun_out/r2x_wide_timing_pf.log gpurun_out/r2x_wide_timing_base.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2x_wide_launches_pf.csv python scripts/wide_timing.py 1000000 2 > gpurun_out/r2x_ncu_pf.log 2>&1
MPINN_B200_LIB=$PWD/multipinn_b200/csrc/_variants/libmpinn_b200_base.so ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2x_wide_launches_base.csv python scripts/wide_timing.py 1000000 2 > gpurun_out/r2x_ncu_base.log 2>&1
python bench.py --workload burgers_1D --steps 5 --warmup 3 > gpurun_out/r2x_bench_burgers_1D.json 2> gpurun_out/r2x_bench_burgers_1D.err
cut -c1-300 gpurun_out/r2x_bench_burgers_1D.json

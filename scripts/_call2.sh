set -x
python scripts/_diag.py > gpurun_out/r2y_diag.log 2>&1; cat gpurun_out/r2y_diag.log | cut -c1-200
python -m pytest tests -m gpu -q -k "not test_full_depth_step_vs_fp64_autograd" > gpurun_out/r2y_gpu_tests.log 2>&1; tail -5 gpurun_out/r2y_gpu_tests.log | cut -c1-300
python -m pytest tests -m gpu -q -s -k "test_full_depth_step_vs_fp64_autograd" > gpurun_out/r2y_gpu_tests_fd.log 2>&1; grep -a "param-grad\|passed\|failed" gpurun_out/r2y_gpu_tests_fd.log | cut -c1-200

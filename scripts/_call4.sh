set -x
python scripts/_diag2.py > gpurun_out/r2v_diag2.log 2>&1; grep "overall\|param 19" gpurun_out/r2v_diag2.log | cut -c1-220
python -m pytest tests -m gpu -q > gpurun_out/r2v_gpu_tests.log 2>&1; tail -4 gpurun_out/r2v_gpu_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; tail -3 gpurun_out/r2v_smoke.log
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; cut -c1-400 gpurun_out/r2v_bench.json
python bench.py --workload burgers_1D --steps 5 --warmup 3 > gpurun_out/r2v_bench_burgers_1D.json 2> gpurun_out/r2v_bench_burgers_1D.err; cut -c1-300 gpurun_out/r2v_bench_burgers_1D.json

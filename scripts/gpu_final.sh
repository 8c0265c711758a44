#!/bin/bash
# Round-end evidence run under gpurun (no profiler): GPU tests, smoke, bench lines for every workload, reference arm.
# usage: scripts/gpu_final.sh <tag>
tag=$1
o=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > $o/${tag}_gpu_tests.log 2>&1; tail -2 $o/${tag}_gpu_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; python scripts/benchline.py $o/${tag}_bench.json
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err; head -c 600 $o/${tag}_bench_reference.json; echo
for w in allen_cahn_2D_1C navier_stokes_2D navier_stokes_3D burgers_1D regression_1D; do
  timeout -s KILL 400 python bench.py --workload $w --no-baselines --steps 5 --warmup 3 > $o/${tag}_bench_$w.json 2> $o/${tag}_bench_$w.err
  echo "$w: $(python scripts/benchline.py $o/${tag}_bench_$w.json | head -1)"
done

#!/bin/bash
# Round-end evidence run under gpurun without the CPU reference arm: GPU tests, smoke, bench lines of every workload,
# ncu launch list (durations only) of the 128-wide engine.  usage: scripts/gpu_final_short.sh <tag>
tag=$1
o=gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > $o/${tag}_gpu_tests.log 2>&1; tail -2 $o/${tag}_gpu_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $o/${tag}_smoke.log 2>&1; tail -1 $o/${tag}_smoke.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > $o/${tag}_bench.json 2> $o/${tag}_bench.err; python scripts/benchline.py $o/${tag}_bench.json
for w in burgers_1D allen_cahn_2D_1C navier_stokes_2D navier_stokes_3D regression_1D; do
  timeout -s KILL 400 python bench.py --workload $w --no-baselines --steps 5 --warmup 3 > $o/${tag}_bench_$w.json 2> $o/${tag}_bench_$w.err
  echo "$w: $(python scripts/benchline.py $o/${tag}_bench_$w.json | head -1)"
done
python scripts/wide_timing.py 4000000 3 > $o/${tag}_wide_timing.log 2>&1; cat $o/${tag}_wide_timing.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $o/${tag}_wide_launches.csv python scripts/wide_timing.py 1000000 2 > $o/${tag}_wide_ncu.log 2>&1

CMD="python bench.py --workload burgers_1D --domain-points 400000 --bound-points 20000 --steps 2 --warmup 3 --no-baselines --no-graphs"
timeout -s KILL 300 $CMD > gpurun_out/plain_ffma.log 2>&1 || { echo plain failed; tail -3 gpurun_out/plain_ffma.log; exit 1; }
timeout -s KILL 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:jet_(fwd|bwd)_ffma<\(int\)2, \(int\)3' -s 4 -c 2 -o gpurun_out/r2_ffma $CMD > gpurun_out/ncu_ffma.log 2>&1
tail -c 200 gpurun_out/ncu_ffma.log

#!/bin/bash
# One `ncu --set full` capture of the domain condition's forward and reverse kernels (run under gpurun).
# usage: scripts/gpu_ncu.sh <out-name> [extra bench.py args]
out=$1; shift
CMD="python bench.py --steps 2 --warmup 3 --no-baselines --no-graphs $*"
$CMD > gpurun_out/plain_$out.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$out.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:jet_(fwd|bwd)_tc<\(int\)[123], \(int\)[23]' -s 4 -c 2 -o gpurun_out/$out $CMD > gpurun_out/ncu_$out.log 2>&1
tail -c 300 gpurun_out/ncu_$out.log
ls -la gpurun_out/

#!/bin/bash
# One `ncu --set full` capture of the layer-wise 128-wide tcgen05 kernels (run under gpurun): the third forward layer,
# wide_last_bwd, the first weight-gradient and the first input-gradient launch of scripts/wide_timing.py (1 M points).
out=${1:-r2_wide}
CMD="python scripts/wide_timing.py 1000000 1"
$CMD > gpurun_out/plain_$out.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$out.log; exit 1; }
cat gpurun_out/plain_$out.log
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:wide_(fwd_layer_kernel|dgrad_layer_kernel|wgrad_kernel|last_bwd)' -s 2 -c 4 -o gpurun_out/$out $CMD > gpurun_out/ncu_$out.log 2>&1
tail -c 300 gpurun_out/ncu_$out.log
ls -la gpurun_out/$out.ncu-rep

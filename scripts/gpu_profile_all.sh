#!/bin/bash
# Round-end evidence run (under gpurun): launch list, one ncu --set full capture of the domain kernels, cycle breakdown.
# usage: scripts/gpu_profile_all.sh <tag>
tag=$1
CMD="python bench.py --steps 2 --warmup 3 --no-baselines --no-graphs"
timeout -s KILL 200 $CMD > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
timeout -s KILL 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:jet_(fwd|bwd)_tc<\(int\)[123], \(int\)[23]' -s 4 -c 2 -o gpurun_out/${tag}_prof $CMD > gpurun_out/ncu_$tag.log 2>&1
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
if [ -f multipinn_b200/csrc/_variants/timing/libmpinn_b200.so ]; then
  MPINN_B200_LIB=multipinn_b200/csrc/_variants/timing/libmpinn_b200.so timeout -s KILL 90 python scripts/tc_timing.py > gpurun_out/${tag}_tc_timing.txt 2>&1
fi
ls -la gpurun_out/ | tail -8

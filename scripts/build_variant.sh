#!/bin/bash
# Build a diagnostic variant of libmpinn_b200.so into multipinn_b200/csrc/_variants/<name>/ (selected at run time with
# MPINN_B200_LIB=<path>): scripts/build_variant.sh timing -DMPINN_TC_TIMING=1
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/multipinn_b200/csrc
out=$src/_variants/$name
mkdir -p $out
for f in api residual jet_ffma jet_tc jet_tc_bwd; do
  ( cd $src && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $f.cu -o $out/$f.o ) &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmpinn_b200.so $out/*.o -lcudart_static -lpthread -ldl -lrt
rm -f $out/*.o
echo built $out/libmpinn_b200.so

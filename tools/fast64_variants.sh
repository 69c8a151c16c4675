#!/bin/bash
# Build variants of the float64 register-resident kernel (its translation unit only, the other objects are compiled once) into tools/_f64_*.so
# for an A/B run on a GPU box:   bash tools/fast64_variants.sh   then   for v in tools/_f64_*.so; do NEMPC_LIB_PATH=$PWD/$v python tools/fast64_time.py; done
set -e
cd "$(dirname "$0")/../pyneuralempc_b200/csrc"
O=/tmp/nempc_objs; mkdir -p $O
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for u in nempc_lib nempc_tc_tu nempc_wide_tu; do [ -f $O/$u.o ] || nvcc $F -c -o $O/$u.o $u.cu & done; wait
i=0
for v in "-DNEMPC_FAST64_JC30=6 -DNEMPC_FAST64_THREADS=128" "-DNEMPC_FAST64_JC30=10 -DNEMPC_FAST64_THREADS=128" "-DNEMPC_FAST64_JC30=15 -DNEMPC_FAST64_THREADS=128" \
         "-DNEMPC_FAST64_JC30=6 -DNEMPC_FAST64_THREADS=96" "-DNEMPC_FAST64_JC30=10 -DNEMPC_FAST64_THREADS=96" "-DNEMPC_FAST64_JC30=10 -DNEMPC_FAST64_THREADS=64" \
         "-DNEMPC_FAST64_JC30=5 -DNEMPC_FAST64_THREADS=128" "-DNEMPC_FAST64_JC30=10 -DNEMPC_FAST64_THREADS=128 -DNEMPC_FAST64_NOSPLIT"; do
  i=$((i+1))
  sp="--split-compile=0"; case "$v" in *NOSPLIT*) sp="";; esac
  ( nvcc $F $sp $v -c -o $O/f64_$i.o nempc_fast64_tu.cu && nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o ../../tools/_f64_$i.so $O/nempc_lib.o $O/nempc_tc_tu.o $O/nempc_wide_tu.o $O/f64_$i.o && echo "$i: $v" ) &
done; wait

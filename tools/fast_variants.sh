#!/bin/bash
# Variants of the register-resident f32 kernel (chunks of the 30-wide second layer, CTAs/SM the register allocator must allow, block size):
# only nempc_lib.cu is recompiled per variant (chunks of 15 neurons are odd: the contracted last stage pairs neurons inside a chunk, so those variants switch it off).   bash tools/fast_variants.sh; for v in tools/_fast_*.so; do NEMPC_LIB_PATH=$PWD/$v python bench.py --steps 50 --no-side-workloads --no-solver --no-cpu-baseline | ...
set -e
cd "$(dirname "$0")/../pyneuralempc_b200/csrc"
O=/tmp/nempc_objs; mkdir -p $O
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
for u in nempc_tc_tu nempc_wide_tu nempc_wide_rt_tu nempc_dmma_tu; do [ -f $O/$u.o ] || nvcc $F -c -o $O/$u.o $u.cu & done
[ -f $O/nempc_fast64_tu.o ] || nvcc $F --split-compile=0 -c -o $O/nempc_fast64_tu.o nempc_fast64_tu.cu &
wait
i=0
for v in "-DNEMPC_FAST_NCHUNK30=3" "-DNEMPC_FAST_NCHUNK30=2 -DNEMPC_FAST_CONTRACT_LAST=0" "-DNEMPC_FAST_NCHUNK30=5" "-DNEMPC_FAST_NCHUNK30=3 -DNEMPC_FAST_THREADS=64" "-DNEMPC_FAST_NCHUNK30=3 -DNEMPC_FAST_THREADS=96" "-DNEMPC_FAST_NCHUNK30=2 -DNEMPC_FAST_CONTRACT_LAST=0 -DNEMPC_FAST_THREADS=64" "-DNEMPC_FAST_NCHUNK30=3 -DNEMPC_FAST_CONTRACT_LAST=0"; do
  i=$((i+1))
  ( nvcc $F $v -c -o $O/lib_$i.o nempc_lib.cu && nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o ../../tools/_fast_$i.so $O/lib_$i.o $O/nempc_tc_tu.o $O/nempc_wide_tu.o $O/nempc_wide_rt_tu.o $O/nempc_dmma_tu.o $O/nempc_fast64_tu.o && echo "$i: $v" ) &
done; wait

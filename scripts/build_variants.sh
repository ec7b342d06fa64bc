#!/bin/bash
# Tuning variants of the specialised RK4 kernel (afast.cu): threads per block / blocks per SM -> lib/libparcels_b200_<tag>.so
#   bash scripts/build_variants.sh b480m1 b384m1 b192m2 b128m3 b256m1      then      PB_LIB=parcels_b200/lib/libparcels_b200_b384m1.so python bench.py ...
# (per-lane shared memory is 480 B: 480 threads fill an SM; 13..16 warps cap the kernel at 128 registers, 12 warps at 168, 8 at 255)
set -e
cd "$(dirname "$0")/../parcels_b200"
python build.py > /dev/null
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -DPB_SMEM_CACHE"
for tag in "$@"; do
  b=${tag#b}; b=${b%m*}; m=${tag#*m}
  nvcc $FLAGS -DPB_BLOCK=$b -DPB_MINBLOCKS=$m -Xptxas -v -c -o lib/obj/afast_$tag.o csrc/afast.cu 2>&1 | grep -A2 "AFastPolicyILi3ELb1EELb0" | grep -E "registers|spill" | tr '\n' ' ' | sed "s/^/[$tag] /; s/$/\n/" &
done
wait
for tag in "$@"; do
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o lib/libparcels_b200_$tag.so lib/obj/engine.o lib/obj/afast_$tag.o lib/obj/agrid.o lib/obj/cgrid.o lib/obj/aslip.o lib/obj/rk45.o lib/obj/advdiff.o lib/obj/hashbuild.o lib/obj/curva.o -ldl
done
ls lib/*.so

#!/bin/bash
# Tuning variants of the specialised RK4 kernels (afast.cu, afast2.cu): threads per block / blocks per SM [/ extra defines]
#   -> lib/libparcels_b200_<tag>.so
#   bash scripts/build_variants.sh b448m1 b416m1 b128m3 b384m1cos0:-DPB_FAST_COS=0
#   then   PB_LIB=parcels_b200/lib/libparcels_b200_b448m1.so python bench.py ...
# (per-lane shared memory is 496 B: 448 threads fill an SM; 13..16 warps cap the kernel at 128 registers, 14 at 144, 12 at 168)
set -e
cd "$(dirname "$0")/../parcels_b200"
python build.py > /dev/null
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -DPB_SMEM_CACHE"
for spec in "$@"; do
  tag=${spec%%:*}; extra=""; [ "$spec" != "$tag" ] && extra=${spec#*:}
  bm=${tag%%cos*}; b=${bm#b}; b=${b%m*}; m=${bm#*m}
  for f in afast afast2; do
    nvcc $FLAGS -DPB_BLOCK=$b -DPB_MINBLOCKS=$m $extra -Xptxas -v -c -o lib/obj/${f}_$tag.o csrc/$f.cu 2>&1 | grep -A2 "Function properties for _Z13advect_kernelI1[12]AFast2\?PolicyILi3ELb1EELb0" | grep -E "registers|spill" | tr '\n' ' ' | sed "s/^/[$tag $f] /; s/$/\n/" &
  done
done
wait
for spec in "$@"; do
  tag=${spec%%:*}
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o lib/libparcels_b200_$tag.so lib/obj/engine.o lib/obj/afast_$tag.o lib/obj/afast2_$tag.o lib/obj/agrid.o lib/obj/cgrid.o lib/obj/aslip.o lib/obj/rk45.o lib/obj/advdiff.o lib/obj/hashbuild.o lib/obj/curva.o -ldl
done
ls lib/*.so

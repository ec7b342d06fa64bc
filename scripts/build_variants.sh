#!/bin/bash
# Build tuning variants of the A-grid kernel (registers/occupancy): lib/libparcels_b200_<tag>.so
set -e
cd "$(dirname "$0")/../parcels_b200"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC"
mkdir -p lib/obj
for tag in "$@"; do
  # tag: mb<N>[s|f]  (s = shared-memory corner cache with float64 copies on float64 grids, f = same cache in the data dtype)
  mb=${tag#mb}; extra=""
  if [[ $mb == *s ]]; then mb=${mb%s}; extra="-DPB_SMEM_CACHE"; fi
  if [[ $mb == *f ]]; then mb=${mb%f}; extra="-DPB_SMEM_CACHE -DPB_SMEM_F32"; fi
  nvcc $FLAGS -DPB_MINBLOCKS=$mb $extra -Xptxas -v -c -o lib/obj/agrid_$tag.o csrc/agrid.cu 2>&1 | grep -A2 "advect_kernelI11AGridPolicyIdfLb1ELi3" | grep -E "registers|spill" | sed "s/^/[$tag] /" &
done
wait
for tag in "$@"; do
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o lib/libparcels_b200_$tag.so lib/obj/engine.o lib/obj/agrid_$tag.o lib/obj/cgrid.o lib/obj/aslip.o lib/obj/rk45.o lib/obj/advdiff.o lib/obj/hashbuild.o
done
ls -la lib/*.so

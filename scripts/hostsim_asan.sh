#!/bin/bash
# Test infrastructure: the host simulation of the kernel sources (oracle/hostsim) rebuilt with AddressSanitizer + UBSan, then a command
# run against it -- out-of-bounds accesses and undefined behaviour in the kernels' index arithmetic show up here without a GPU.
#   bash scripts/hostsim_asan.sh python scripts/fuzz_hostsim.py 80 5
#   bash scripts/hostsim_asan.sh python -m pytest tests -m gpu -q -x -p no:cacheprovider --deselect tests/test_gpu_decomposed.py
cd "$(dirname "$0")/.."
python -c "from oracle.hostsim import build as hb; hb.build()" || exit 1   # generates oracle/_build/hostsim/src
out=${HS_ASAN_DIR:-/tmp/parcels_b200_hostsim_asan}   # (outside the tree: the snapshot that travels to the GPU box stays small)
mkdir -p $out
if [ ! -f $out/libhs_asan.so ] || [ oracle/_build/hostsim/libparcels_b200_hostsim.so -nt $out/libhs_asan.so ]; then
  for f in oracle/_build/hostsim/src/*.cpp; do
    g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -w -fsanitize=address,undefined -fno-sanitize-recover=undefined \
        -I oracle/hostsim/include -DPB_SMEM_CACHE -DPB_MINBLOCKS=4 -c -o $out/$(basename ${f%.cpp}).o $f &
  done
  wait
  g++ -shared -fsanitize=address,undefined -o $out/libhs_asan.so $out/*.o || exit 1
fi
PB_LIB=$out/libhs_asan.so PB_HOSTSIM_TEST=1 PYTHONPATH=$PWD ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
  LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) "$@"

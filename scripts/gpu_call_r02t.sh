#!/bin/bash
# r02t: raw block as node records (no transposition at a refill, T-lerp two corners at a time) against the previous layout
tag=${1:-r02t}
out=gpurun_out
mkdir -p $out
( time python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_parity.py -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -2 $out/${tag}_tests.log
for w in c2 ns c4; do
  st=4; [ $w = c2 ] && st=15
  for v in prev new prev new; do
    lib=parcels_b200/lib/libparcels_b200.so; [ $v = prev ] && lib=parcels_b200/lib/libparcels_b200_prev.so
    PB_LIB=$PWD/$lib python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_${v}_$w.json 2>> $out/${tag}_sweep.err
    python scripts/bench_summary.py --brief "$v $w" $out/${tag}_${v}_$w.json
  done
done

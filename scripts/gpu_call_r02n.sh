#!/bin/bash
# r02n: afast2 (two-stage body, inline side paths) with the diffusion increment out of line against afast.cu on c4; 128 x 3 blocks;
# then the full suite and the driver's default bench line of this tree
tag=${1:-r02n}
out=gpurun_out
mkdir -p $out
for v in 1 2 1 2; do
  PB_FAST_KERNEL=$v python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_v${v}_c4.json 2>> $out/${tag}_sweep.err
  python scripts/bench_summary.py --brief "v$v c4" $out/${tag}_v${v}_c4.json
done
for w in c2 ns; do
  st=4; [ $w = c2 ] && st=15
  PB_LIB=$PWD/parcels_b200/lib/libparcels_b200_b128.so python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_b128_$w.json 2>> $out/${tag}_sweep.err
  python scripts/bench_summary.py --brief "afast2 128x3 $w" $out/${tag}_b128_$w.json
  python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_default_$w.json 2>> $out/${tag}_sweep.err
  python scripts/bench_summary.py --brief "afast2 384x1 $w" $out/${tag}_default_$w.json
done
( time python -m pytest tests -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
( time python bench.py ) > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
tail -3 $out/${tag}_bench_default.err
python scripts/bench_summary.py $out/${tag}_bench_default.json
PB_FAST_KERNEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 1 -c 1 -o $out/${tag}_advect_v2_c4 -f \
    python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_c4.log 2>&1
python scripts/ncu_summary.py $out/${tag}_advect_v2_c4.ncu-rep > $out/${tag}_ncu_summary_v2_c4.txt 2>&1
grep -E "time_dur|inst_executed.sum|issue_active|registers_per|stalls|SASS" $out/${tag}_ncu_summary_v2_c4.txt

#!/bin/bash
# r02l: afast2 as a two-stage loop body (PB_FAST_KERNEL=2) against afast.cu on c2 / ns / c4; GPU tests of the trees' newest parts
tag=${1:-r02l}
out=gpurun_out
mkdir -p $out
( time python -m pytest tests/test_gpu_advdiff.py tests/test_gpu_fast_kernel.py tests/test_gpu_multigrid.py tests/test_gpu_scalar.py -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
for w in c2 ns c4; do
  st=4; [ $w = c2 ] && st=15
  for v in 1 2 1 2; do
    PB_FAST_KERNEL=$v python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_v${v}_$w.json 2>> $out/${tag}_sweep.err
    python scripts/bench_summary.py --brief "v$v $w" $out/${tag}_v${v}_$w.json
  done
done
for v in default cginl cg4 default cginl; do
  lib=parcels_b200/lib/libparcels_b200_$v.so; [ $v = default ] && lib=parcels_b200/lib/libparcels_b200.so
  PB_LIB=$PWD/$lib python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_c3_$v.json 2>> $out/${tag}_sweep.err
  python scripts/bench_summary.py --brief "c3 $v" $out/${tag}_c3_$v.json
done
PB_LIB=$PWD/parcels_b200/lib/libparcels_b200_cginl.so python bench.py --workload c3_3d --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_c3_3d_cginl.json 2>> $out/${tag}_sweep.err
python scripts/bench_summary.py --brief "c3_3d cginl" $out/${tag}_c3_3d_cginl.json
python bench.py --workload c3_3d --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_c3_3d_default.json 2>> $out/${tag}_sweep.err
python scripts/bench_summary.py --brief "c3_3d default" $out/${tag}_c3_3d_default.json
PB_FAST_KERNEL=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 3 -c 1 -o $out/${tag}_advect_v2_c2 -f \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_c2.log 2>&1
python scripts/ncu_summary.py $out/${tag}_advect_v2_c2.ncu-rep > $out/${tag}_ncu_summary_v2_c2.txt 2>&1
grep -E "time_dur|inst_executed.sum|issue_active|registers_per|stalls|SASS" $out/${tag}_ncu_summary_v2_c2.txt

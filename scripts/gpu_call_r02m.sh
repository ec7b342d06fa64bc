#!/bin/bash
# r02m: afast2 (two-stage loop body) with the side path INLINE at both sites, against afast.cu and the out-of-line form
tag=${1:-r02m}
out=gpurun_out
mkdir -p $out
for w in c2 ns c4; do
  st=4; [ $w = c2 ] && st=15
  PB_FAST_KERNEL=1 python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_v1_$w.json 2>> $out/${tag}_sweep.err
  python scripts/bench_summary.py --brief "v1 $w" $out/${tag}_v1_$w.json
  PB_FAST_KERNEL=2 PB_LIB=$PWD/parcels_b200/lib/libparcels_b200_sideinl.so python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_v2inl_$w.json 2>> $out/${tag}_sweep.err
  python scripts/bench_summary.py --brief "v2 side inline $w" $out/${tag}_v2inl_$w.json
done
python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_c3.json 2>> $out/${tag}_sweep.err
python scripts/bench_summary.py --brief "c3 default (inline hash, advection-only instantiation)" $out/${tag}_c3.json
PB_FAST_KERNEL=2 PB_LIB=$PWD/parcels_b200/lib/libparcels_b200_sideinl.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 1 -c 1 -o $out/${tag}_advect_v2inl_ns -f \
    python bench.py --workload ns --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_ns.log 2>&1
python scripts/ncu_summary.py $out/${tag}_advect_v2inl_ns.ncu-rep > $out/${tag}_ncu_summary_v2inl_ns.txt 2>&1
grep -E "time_dur|inst_executed.sum|issue_active|registers_per|stalls|SASS" $out/${tag}_ncu_summary_v2inl_ns.txt

#!/bin/bash
# r02h: the round's candidate tree -- full GPU suite; the driver's default bench line; curvilinear kernel with the custom sin / cos
# (register variants), its end-to-end arm with the in-place compacted download; float32 Box-Muller (c4); the row f-4 workloads
# (3-D curvilinear at ORCA025 size, 2-D at ORCA12 size) with parity samples; ncu --set full of c3 / c4 / ns; launch list.
tag=${1:-r02h}
out=gpurun_out
mkdir -p $out
( time python -m pytest tests -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
( time python bench.py ) > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
tail -3 $out/${tag}_bench_default.err
python scripts/bench_summary.py $out/${tag}_bench_default.json
run() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --extras "" > $out/${tag}_${label}.json 2>> $out/${tag}_sweep.err
  python scripts/bench_summary.py --brief "$label" $out/${tag}_${label}.json
}
for v in cg3 cg5; do
  run c3_$v PB_LIB=$PWD/parcels_b200/lib/libparcels_b200_$v.so -- --workload c3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e
done
run c3_3d X=1 -- --workload c3_3d --steps 4 --warmup 3
python scripts/bench_summary.py $out/${tag}_c3_3d.json
run c3_3d_cg3 PB_LIB=$PWD/parcels_b200/lib/libparcels_b200_cg3.so -- --workload c3_3d --steps 4 --warmup 3 --no-cpu-baseline --no-e2e
run c3_orca12 X=1 -- --workload c3_orca12 --steps 4 --warmup 3
python scripts/bench_summary.py $out/${tag}_c3_orca12.json
for w in c3 c4 ns; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 1 -c 1 -o $out/${tag}_advect_$w -f \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_$w.log 2>&1
  python scripts/ncu_summary.py $out/${tag}_advect_$w.ncu-rep > $out/${tag}_ncu_summary_$w.txt 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_launches_default.log 2>&1
ls -la $out/${tag}*.ncu-rep

#!/bin/bash
# r02k: tree with constant-memory polynomial coefficients, 168-register curvilinear kernels, host-side batch-flag fast paths,
# multi-grid fields, XLinear scalars on curvilinear grids: full suite, default bench line, dense-release workload, ncu of c3 / c2.
tag=${1:-r02k}
out=gpurun_out
mkdir -p $out
( time python -m pytest tests -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
( time python bench.py ) > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
tail -3 $out/${tag}_bench_default.err
python scripts/bench_summary.py $out/${tag}_bench_default.json
python bench.py --workload ns_dense --steps 5 --warmup 3 --extras "" > $out/${tag}_ns_dense.json 2>> $out/${tag}_sweep.err
python scripts/bench_summary.py $out/${tag}_ns_dense.json
python bench.py --workload ns --sorted --steps 5 --warmup 3 --extras "" --no-cpu-baseline --no-e2e > $out/${tag}_ns_sorted.json 2>> $out/${tag}_sweep.err
python scripts/bench_summary.py --brief "ns sorted" $out/${tag}_ns_sorted.json
for w in c3 c2 ns_dense; do
  s=1; [ $w = c2 ] && s=3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s $s -c 1 -o $out/${tag}_advect_$w -f \
      python bench.py --workload $w --steps 2 --warmup $s --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_$w.log 2>&1
  python scripts/ncu_summary.py $out/${tag}_advect_$w.ncu-rep > $out/${tag}_ncu_summary_$w.txt 2>&1
done
( time python bench.py --impl reference ) > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -2 $out/${tag}_bench_reference.json | cut -c1-600
ls -la $out/${tag}*.ncu-rep

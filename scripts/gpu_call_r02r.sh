#!/bin/bash
# r02r: evidence of the final tree (one afast.cu with two schedules, host-side execute() passes trimmed) -- full GPU suite, the driver's
# kernels (ns, c2: afast2 128 x 3; c3: curvilinear advection-only instantiation; c4: afast.cu with diffusion)
tag=${1:-r02r}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/${tag}_smi.txt
( time python -m pytest tests -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
( time python bench.py --impl reference ) > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
( time python bench.py ) > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
tail -3 $out/${tag}_bench_default.err
python scripts/bench_summary.py $out/${tag}_bench_default.json $out/${tag}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_launches_default.log 2>&1
for w in ns c2 c4; do
  s=1; [ $w = c2 ] && s=3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s $s -c 1 -o $out/${tag}_advect_$w -f \
      python bench.py --workload $w --steps 2 --warmup $s --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_$w.log 2>&1
  python scripts/ncu_summary.py $out/${tag}_advect_$w.ncu-rep > $out/${tag}_ncu_summary_$w.txt 2>&1
done
ls -la $out/${tag}*.ncu-rep

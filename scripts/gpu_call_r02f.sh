#!/bin/bash
# r02f: first consolidated B200 call of round 2 after the container was re-created (earlier r02b-e outputs were lost):
# whole GPU suite, the driver's two bench invocations, launch list, block-size sweep of the specialised RK4 kernel,
# ncu --set full of the RK4 kernel (c2, ns), the fused-diffusion kernel (c4) and the curvilinear kernel (c3).
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/${tag}_smi.txt
( time python -m pytest tests -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
( time python bench.py ) > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
tail -3 $out/${tag}_bench_default.err
( time python bench.py --impl reference ) > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -3 $out/${tag}_bench_reference.err
python scripts/bench_summary.py $out/${tag}_bench_default.json $out/${tag}_bench_reference.json
# launch list of the default command (bounded: 2 steps, no CPU legs)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_launches_default.log 2>&1
# block-size sweep
for v in default b480m1 b256m1 b192m2 b128m3 b96m4; do
  lib=parcels_b200/lib/libparcels_b200_$v.so; [ $v = default ] && lib=parcels_b200/lib/libparcels_b200.so
  [ -f $lib ] || continue
  for w in c2 ns; do
    st=4; [ $w = c2 ] && st=15
    PB_LIB=$PWD/$lib python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_sweep_${v}_${w}.json 2>> $out/${tag}_sweep.err
    python scripts/bench_summary.py --brief "$v $w" $out/${tag}_sweep_${v}_${w}.json
  done
done
# generic kernel for A/B
for w in c2 ns; do
  PB_DISABLE_FAST_KERNEL=1 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_generic_${w}.json 2>> $out/${tag}_sweep.err
  python scripts/bench_summary.py --brief "generic $w" $out/${tag}_generic_${w}.json
done
# e2e: pipelined chunks vs three-call path
for pc in 0 4; do
  for w in c2 ns c3; do
    python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --pipeline $pc --extras "" > $out/${tag}_pipe${pc}_${w}.json 2>> $out/${tag}_sweep.err
    python scripts/bench_summary.py --brief "pipeline=$pc $w" $out/${tag}_pipe${pc}_${w}.json
  done
done
for w in c2 ns c3 c4; do
  s=1; [ $w = c2 ] && s=3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s $s -c 1 -o $out/${tag}_advect_$w -f \
      python bench.py --workload $w --steps 2 --warmup $s --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_$w.log 2>&1
  python scripts/ncu_summary.py $out/${tag}_advect_$w.ncu-rep > $out/${tag}_ncu_summary_$w.txt 2>&1
done
ls -la $out/*.ncu-rep

#!/bin/bash
# r02g: afast2.cu (stages written out, out-of-line side path, polynomial cos) against afast.cu and block / register variants; full
# GPU suite of the new tree (curvilinear XLinear, in-kernel migration incl. CUDA IPC between two processes, ulp eval tolerances);
# ncu --set full of the new kernel on c2 and ns; compute-sanitizer memcheck / racecheck of the small workloads.
tag=${1:-r02g}
out=gpurun_out
mkdir -p $out
( time python -m pytest tests -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
run() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py "$@" --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_${label}.json 2>> $out/${tag}_sweep.err
  python scripts/bench_summary.py --brief "$label" $out/${tag}_${label}.json
}
for w in c2 ns; do
  st=4; [ $w = c2 ] && st=15
  run v2_default_$w PB_FAST_KERNEL=2 -- --workload $w --steps $st --warmup 3
  run v1_default_$w PB_FAST_KERNEL=1 -- --workload $w --steps $st --warmup 3
  for v in b416m1r152 b448m1r144 b128m3 b192m2 b384m1cos0; do
    run v2_${v}_$w PB_FAST_KERNEL=2 PB_LIB=$PWD/parcels_b200/lib/libparcels_b200_$v.so -- --workload $w --steps $st --warmup 3
  done
done
run v2_default_c4 PB_FAST_KERNEL=2 -- --workload c4 --steps 4 --warmup 3
run v1_default_c4 PB_FAST_KERNEL=1 -- --workload c4 --steps 4 --warmup 3
run v2_b416m1r152_c4 PB_FAST_KERNEL=2 PB_LIB=$PWD/parcels_b200/lib/libparcels_b200_b416m1r152.so -- --workload c4 --steps 4 --warmup 3
for w in c2 ns c4; do
  s=1; [ $w = c2 ] && s=3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s $s -c 1 -o $out/${tag}_advect_$w -f \
      python bench.py --workload $w --steps 2 --warmup $s --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_$w.log 2>&1
  python scripts/ncu_summary.py $out/${tag}_advect_$w.ncu-rep > $out/${tag}_ncu_summary_$w.txt 2>&1
done
for w in c2_small c3_small c4_small; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python bench.py --workload $w --particles 20000 --steps 1 --warmup 1 --no-cpu-baseline --extras "" \
      > $out/${tag}_memcheck_$w.log 2>&1; echo "memcheck $w rc=$?" | tee -a $out/${tag}_sanitizer.txt
  grep -E "ERROR SUMMARY|Invalid|error" $out/${tag}_memcheck_$w.log | head -5 | tee -a $out/${tag}_sanitizer.txt
done
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python bench.py --workload c2_small --particles 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --extras "" \
    > $out/${tag}_racecheck_c2_small.log 2>&1; echo "racecheck c2_small rc=$?" | tee -a $out/${tag}_sanitizer.txt
grep -E "RACECHECK SUMMARY|hazard" $out/${tag}_racecheck_c2_small.log | head -5 | tee -a $out/${tag}_sanitizer.txt
ls -la $out/${tag}*.ncu-rep

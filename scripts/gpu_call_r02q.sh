#!/bin/bash
# r02q: afast.cu after the two schedules were unified into one policy (SCHED 1 / 2): bit-for-bit tests and timings of c2 / ns / c4
tag=${1:-r02q}
out=gpurun_out
mkdir -p $out
( time python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_parity.py tests/test_gpu_decomposed.py -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
for w in c2 ns c4; do
  st=4; [ $w = c2 ] && st=15
  for v in 1 2; do
    PB_FAST_KERNEL=$v python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_s${v}_$w.json 2>> $out/${tag}_sweep.err
    python scripts/bench_summary.py --brief "schedule $v $w" $out/${tag}_s${v}_$w.json
  done
done
python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline --extras "" > $out/${tag}_default_c4.json 2>> $out/${tag}_sweep.err
python scripts/bench_summary.py --brief "default c4" $out/${tag}_default_c4.json

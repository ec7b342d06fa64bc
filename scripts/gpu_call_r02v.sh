#!/bin/bash
# r02v: ncu --set full of the committed RK4 kernel (raw block as node records) on ns and c2
tag=${1:-r02v}
out=gpurun_out
mkdir -p $out
for w in ns c2; do
  s=1; [ $w = c2 ] && s=3
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s $s -c 1 -o $out/${tag}_advect_$w -f \
      python bench.py --workload $w --steps 2 --warmup $s --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_$w.log 2>&1
done
ls -la $out/${tag}*.ncu-rep

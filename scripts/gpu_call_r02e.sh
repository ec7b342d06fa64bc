#!/bin/bash
# r02e: block-size / register sweep of the specialised RK4 kernel (cells in shared memory), ncu of the c3 and c4 kernels at full size
tag=${1:-r02e}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_parity.py -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -2 $out/${tag}_tests.log
for v in default b480m1 b192m2 b128m3 b256m1 b96m4; do
  lib=parcels_b200/lib/libparcels_b200_$v.so; [ $v = default ] && lib=parcels_b200/lib/libparcels_b200.so
  for w in c2 ns c4; do
    st=4; [ $w = c2 ] && st=15
    PB_LIB=$PWD/$lib python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_sweep_${v}_${w}.json 2>> $out/${tag}_sweep.err
    python - "$out/${tag}_sweep_${v}_${w}.json" "$v $w" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); m = d['measured']
    print(f"{sys.argv[2]}: value {d['value']:.3e} kernel_ms {m['kernel_ms_per_launch']:.2f}")
except Exception as e:
    print(sys.argv[2], "unreadable", e)
PY
  done
done
ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 3 -c 1 -o $out/${tag}_advect_c2 -f \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 1 -c 1 -o $out/${tag}_advect_c3 -f \
    python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 1 -c 1 -o $out/${tag}_advect_c4 -f \
    python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_c4.log 2>&1
ls -la $out/${tag}_advect_*.ncu-rep

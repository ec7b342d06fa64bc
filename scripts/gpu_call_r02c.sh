#!/bin/bash
# r02c: cached-reciprocal divisions in the specialised RK4 kernel: parity + timing + ncu
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_ownership.py tests/test_gpu_parity.py tests/test_gpu_decomposed.py -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
for w in c2 ns c4; do
  st=5; [ $w = c2 ] && st=20
  python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_bench_${w}.json 2> $out/${tag}_bench_${w}.err
done
ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 3 -c 1 -o $out/${tag}_advect_c2 -f \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_c2.log 2>&1
for f in $out/${tag}_bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    m = d['measured']
    print(f"value {d['value']:.3e}  kernel_ms {m['kernel_ms_per_launch']:.2f} refills {m['corner_cache_refills_per_launch']} deleted {m['deleted_per_launch']} variant {m['kernel_variant']}")
except Exception as e:
    print("unreadable:", e)
PY
)"; done
tail -5 $out/${tag}_bench_ns.err

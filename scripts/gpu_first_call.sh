#!/bin/bash
# One gpurun call that re-establishes the measured state of the tree on a B200 (about 12 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_first_call.sh r02a'
# Everything lands under gpurun_out/<tag>_*; copy what should be judged into profiles/.
# 1. the GPU parity suite; 2. default bench line (config 2) and the pipelined end-to-end experiment; 3. north-star, config 3, config 4 lines;
# 4. ncu launch list + one full capture of the default workload's kernel.
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_gpu_tests.log
tail -3 $out/${tag}_gpu_tests.log
python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; tail -c 600 $out/${tag}_bench_c2.json
for c in 2 4 8; do
  python bench.py --no-cpu-baseline --steps 10 --pipeline $c > $out/${tag}_bench_c2_pipeline$c.json 2>> $out/${tag}_bench_c2.err
done
python bench.py --workload ns --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_ns.json 2> $out/${tag}_bench_ns.err
python bench.py --workload ns --steps 3 --warmup 3 --no-cpu-baseline --pipeline 8 > $out/${tag}_bench_ns_pipeline8.json 2>> $out/${tag}_bench_ns.err
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 3 -c 1 -o $out/${tag}_advect_c2 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_ncu_full.log 2>&1
for f in $out/${tag}_bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"value {d['value']:.3e}  e2e {d['e2e']['value'] if d.get('e2e') else None}  frac {d['roofline']['frac']:.3f}  kernel_ms {d['config']['kernel_ms_per_launch']:.2f}")
except Exception as e:
    print("unreadable:", e)
PY
)"; done

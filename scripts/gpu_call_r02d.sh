#!/bin/bash
# r02d: full new bench line (ns + extras c2/c3/c4 + parity samples), reference arm, full-size parity tests, ncu of the RK4 kernel
tag=${1:-r02d}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_fast_kernel.py tests/test_gpu_ownership.py tests/test_gpu_parity.py tests/test_gpu_decomposed.py tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
( time python bench.py --steps 10 --warmup 3 ) > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
tail -3 $out/${tag}_bench_default.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
tail -3 $out/${tag}_bench_reference.err
ncu --set full --clock-control none --import-source on -k regex:advect_kernel -s 3 -c 1 -o $out/${tag}_advect_c2 -f \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --extras "" > $out/${tag}_ncu_c2.log 2>&1
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02d_bench_default.json").read().strip().splitlines()[-1])
    def show(name, b):
        m = b["measured"]; e = b.get("e2e") or {}
        p = b.get("parity_sample") or {}
        print(f"{name}: value {b['value']:.3e} e2e {e.get('value', 0):.3e} kernel_ms {m['kernel_ms_per_launch']:.2f} variant {m['kernel_variant']} frac {b['roofline']['frac']:.3f} "
              f"parity ok={p.get('ok')} max_ulp={p.get('max_ulp')} ei_mm={p.get('ei_mismatch')} deleted {p.get('deleted_gpu')}/{p.get('deleted_oracle')} cpu {b.get('cpu_baseline', {}).get('value', 0):.3e}")
    show("ns", d)
    for k, v in d.get("extra", {}).items():
        show(k, v)
    r = json.loads(open("gpurun_out/r02d_bench_reference.json").read().strip().splitlines()[-1])
    print("reference:", r.get("value"), r["cpu_baseline"]["sample"] if "cpu_baseline" in r else r)
    print("same config:", d["config"] == r.get("config"))
except Exception as ex:
    print("summary failed:", ex)
PY

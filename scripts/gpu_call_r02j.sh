#!/bin/bash
# r02j (2 GPUs): mode D with in-kernel migration over peer memory (CUDA IPC + NVLink) -- the decomposed tests incl. NCCL, the driver's
# multi-GPU bench invocation at N = 2 (mode R line + mode_d block + bit-exactness check), and the collective transport for comparison.
tag=${1:-r02j}
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
( time python -m pytest tests/test_gpu_decomposed.py -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 5 --warmup 3 ) \
    > $out/${tag}_bench_2gpu.json 2> $out/${tag}_bench_2gpu.err
tail -3 $out/${tag}_bench_2gpu.err
python scripts/bench_summary.py $out/${tag}_bench_2gpu.json
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --workload c5 --steps 3 --warmup 2 \
    --mode-d-transport collective ) > $out/${tag}_bench_c5_collective.json 2> $out/${tag}_bench_c5_collective.err
tail -3 $out/${tag}_bench_c5_collective.err
python - <<'PY'
import json
for f in ("gpurun_out/r02j_bench_2gpu.json", "gpurun_out/r02j_bench_c5_collective.json"):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        md = d.get("mode_d") or {}
        print(f, "value", d.get("value"), "| mode_d value", md.get("value"), "transport", md.get("transport"), "rounds", md.get("advect_rounds_per_pass"),
              "migr/pass", md.get("migrations_per_pass"), "kernel ms", md.get("kernel_ms_per_pass_max_rank"), "coll ms", md.get("collective_ms_per_pass_max_rank"),
              "ms/pass", md.get("ms_per_step"), "bitexact", md.get("bitexact_check"))
    except Exception as e:
        print(f, "unreadable", e)
PY

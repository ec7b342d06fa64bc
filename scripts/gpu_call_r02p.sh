#!/bin/bash
# r02p (4 GPUs): the driver's multi-GPU bench invocation at N = 4 -- mode R line + mode_d block (config 5 over peer memory) + bit-exactness check
tag=${1:-r02p}
out=gpurun_out
mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 4 --steps 5 --warmup 3 ) \
    > $out/${tag}_bench_4gpu.json 2> $out/${tag}_bench_4gpu.err
tail -3 $out/${tag}_bench_4gpu.err
python scripts/bench_summary.py $out/${tag}_bench_4gpu.json | cut -c1-1500

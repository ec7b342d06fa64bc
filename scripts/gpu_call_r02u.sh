#!/bin/bash
# r02u: the committed tree at the end of round 2: full GPU suite + the driver's default bench line
tag=${1:-r02u}
out=gpurun_out
mkdir -p $out
( time python -m pytest tests -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
( time python bench.py ) > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err
tail -3 $out/${tag}_bench_default.err
python scripts/bench_summary.py $out/${tag}_bench_default.json

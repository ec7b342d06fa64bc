#!/bin/bash
# r02s: last check of the committed tree: full GPU suite, smoke(), the c4 end-to-end arm after the strided ei helpers
tag=${1:-r02s}
out=gpurun_out
mkdir -p $out
( time python -m pytest tests -m gpu -x -q ) > $out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --extras "" > $out/${tag}_c4.json 2>> $out/${tag}_sweep.err
python scripts/bench_summary.py --brief "c4" $out/${tag}_c4.json

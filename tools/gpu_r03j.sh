#!/bin/bash
# Round 2 (third session): after the parallel node norms and the spines of nondeterministic prediction.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r03j_pytest.txt 2>&1; tail -3 $OUT/r03j_pytest.txt
timeout 900 python bench.py --steps 3 > $OUT/r03j_bench.json 2> $OUT/r03j_bench.err; tail -c 200 $OUT/r03j_bench.err; head -c 300 $OUT/r03j_bench.json; echo

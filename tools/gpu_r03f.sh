#!/bin/bash
# Round 2 (third session): GPU suite with the nondeterministic-prediction tests, then the default bench line.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r03f_pytest.txt 2>&1; tail -3 $OUT/r03f_pytest.txt
timeout 900 python bench.py > $OUT/r03f_bench.json 2> $OUT/r03f_bench.err; tail -c 300 $OUT/r03f_bench.err; head -c 700 $OUT/r03f_bench.json; echo

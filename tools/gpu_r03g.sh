#!/bin/bash
# Round 2 (third session): GPU suite with the colour-sequence and nondeterministic-prediction tests, bench line.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv > $OUT/r03g_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r03g_pytest.txt 2>&1; tail -3 $OUT/r03g_pytest.txt
timeout 900 python bench.py > $OUT/r03g_bench.json 2> $OUT/r03g_bench.err; tail -c 300 $OUT/r03g_bench.err; head -c 400 $OUT/r03g_bench.json; echo

#!/bin/bash
# Round 2 (third session), closing run: GPU suite with the colour-sequence and nondeterministic-prediction tests, bench line.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv > $OUT/r03i_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/r03i_pytest.txt 2>&1; tail -3 $OUT/r03i_pytest.txt
timeout 900 python bench.py > $OUT/r03i_bench.json 2> $OUT/r03i_bench.err; tail -c 300 $OUT/r03i_bench.err; head -c 400 $OUT/r03i_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r03i_bench_reference.json 2>> $OUT/r03i_bench.err; head -c 300 $OUT/r03i_bench_reference.json; echo

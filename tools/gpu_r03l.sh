#!/bin/bash
# Round 2 (third session): ncu of the two new launch kinds: a colour P frame (cluster of 8) and a still with
# nondeterministic prediction.
set -u
OUT=gpurun_out
mkdir -p $OUT
python tools/coder_once.py nd512 > $OUT/r03l_runs.txt 2>&1
python tools/coder_once.py colour 3 >> $OUT/r03l_runs.txt 2>&1; cat $OUT/r03l_runs.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -c 1 \
    -f -o $OUT/r03l_nd512 python tools/coder_once.py nd512 > $OUT/r03l_nd512_ncu.log 2>&1; tail -1 $OUT/r03l_nd512_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -s 1 -c 1 \
    -f -o $OUT/r03l_colour_p python tools/coder_once.py colour 2 > $OUT/r03l_colour_p_ncu.log 2>&1; tail -1 $OUT/r03l_colour_p_ncu.log

#!/bin/bash
# Round 2 (third session), ncu of HEAD: launch list of the bench, the full-batch launch (traffic), the clustered
# single-frame launch, a clustered P-frame launch.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/r03h_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-configs > $OUT/r03h_bench_under_ncu.log 2>&1
grep -c fiasco_tile_kernel $OUT/r03h_bench_launches.csv
bash tools/prof_batch.sh r03h
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -c 1 \
    -f -o $OUT/r03h_single python tools/gpu_check.py big > $OUT/r03h_single_ncu.log 2>&1; tail -1 $OUT/r03h_single_ncu.log
FBQ_FRAMES=8 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -s 1 -c 1 \
    -f -o $OUT/r03h_video python tools/video_quick.py 1 > $OUT/r03h_video_ncu.log 2>&1; tail -1 $OUT/r03h_video_ncu.log
ls -la $OUT | grep r03h

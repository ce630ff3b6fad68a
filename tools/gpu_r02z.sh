#!/bin/bash
# Round 2, closing evidence run on one B200: the GPU suite, bench.py (full line), ncu of HEAD: launch list of
# the bench, one clustered single-frame launch, one clustered P-frame launch.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv > $OUT/r02z_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02z_pytest.txt 2>&1; tail -2 $OUT/r02z_pytest.txt
python bench.py --steps 5 --warmup 3 > $OUT/r02z_bench.json 2> $OUT/r02z_bench.err; tail -c 400 $OUT/r02z_bench.err; head -c 600 $OUT/r02z_bench.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r02z_bench_reference.json 2>> $OUT/r02z_bench.err; head -c 400 $OUT/r02z_bench_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/r02z_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-configs > $OUT/r02z_bench_under_ncu.log 2>&1
grep -c fiasco_tile_kernel $OUT/r02z_bench_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -c 1 \
    -f -o $OUT/r02z_single python tools/gpu_check.py big > $OUT/r02z_single_ncu.log 2>&1; tail -1 $OUT/r02z_single_ncu.log
FBQ_FRAMES=8 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -s 1 -c 1 \
    -f -o $OUT/r02z_video python tools/video_quick.py 1 > $OUT/r02z_video_ncu.log 2>&1; tail -1 $OUT/r02z_video_ncu.log

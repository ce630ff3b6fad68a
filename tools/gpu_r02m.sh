#!/bin/bash
# Round 2 evidence run: the whole GPU suite (cluster tests included), config 5 after the motion-search
# reductions, and the ncu captures of HEAD: launch list of bench.py, full-batch launch (traffic), one
# clustered single-frame launch, one P-frame launch.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv > $OUT/r02m_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/r02m_pytest.txt 2>&1; tail -3 $OUT/r02m_pytest.txt
timeout 600 python tools/video_quick.py 2 e9d88f99690abf5b88c449478ff1dbf3 > $OUT/r02m_video_c5.txt 2>&1; cat $OUT/r02m_video_c5.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r02m_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-configs > $OUT/r02m_bench_under_ncu.log 2>&1
grep -c fiasco_tile_kernel $OUT/r02m_bench_launches.csv
bash tools/prof_batch.sh r02m
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -c 1 \
    -f -o $OUT/r02m_single python tools/gpu_check.py big > $OUT/r02m_single_ncu.log 2>&1; tail -2 $OUT/r02m_single_ncu.log
FBQ_FRAMES=8 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -s 1 -c 1 \
    -f -o $OUT/r02m_video python tools/video_quick.py 1 > $OUT/r02m_video_ncu.log 2>&1; tail -2 $OUT/r02m_video_ncu.log
ls -la $OUT | tail -20

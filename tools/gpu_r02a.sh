#!/bin/bash
# Round 2, first GPU call: where does the time of ONE stream go (the fiasco_coder() single-image case and
# the P-frame kernel)?  Lap timers of the diagnostics build, one full ncu capture of each.  Output in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02a.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv > $OUT/r02a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02a_pytest.txt 2>&1; tail -5 $OUT/r02a_pytest.txt
# 1. one 1024^2 frame, lap shares per thread-block shape (diagnostics build)
for nt in 512 256 128; do
  FB200_LIB=gpurun_exp/laps/libfiasco_b200.so FB200_NT=$nt timeout 300 python tools/gpu_check.py big \
      > $OUT/r02a_single_laps_nt$nt.txt 2>&1
  tail -3 $OUT/r02a_single_laps_nt$nt.txt
done
# 2. product build, the same frame: kernel time, then the full capture with source-level samples
timeout 300 python tools/gpu_check.py big > $OUT/r02a_single_product.txt 2>&1
tail -2 $OUT/r02a_single_product.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -c 1 \
    -f -o $OUT/r02a_single python tools/gpu_check.py big > $OUT/r02a_single_ncu.log 2>&1
tail -2 $OUT/r02a_single_ncu.log
# 3. the motion kernel: config 5 timing (torch-free), launch list, one full capture of a P-frame launch
timeout 600 python tools/video_quick.py 2 e9d88f99690abf5b88c449478ff1dbf3 > $OUT/r02a_video_c5.txt 2>&1
cat $OUT/r02a_video_c5.txt
FBQ_FRAMES=8 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file $OUT/r02a_video_launches.csv python tools/video_quick.py 1 > $OUT/r02a_video_launches.log 2>&1
FBQ_FRAMES=8 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fiasco_tile_kernel -s 1 -c 1 \
    -f -o $OUT/r02a_video python tools/video_quick.py 1 > $OUT/r02a_video_ncu.log 2>&1
tail -2 $OUT/r02a_video_ncu.log
ls -la $OUT

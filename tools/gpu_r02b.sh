#!/bin/bash
# Round 2, cluster per stream: parity on hardware, single-frame latency per cluster size, tile-split configs.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_event_reasons.active --format=csv > $OUT/r02b_smi.txt
for c in 1 2 4 8; do
  FB200_CLUSTER=$c timeout 300 python tools/gpu_check.py big > $OUT/r02b_single_c$c.txt 2>&1
  tail -3 $OUT/r02b_single_c$c.txt | head -1
done
timeout 300 python tools/gpu_check.py big > $OUT/r02b_single_auto.txt 2>&1; tail -3 $OUT/r02b_single_auto.txt | head -1
FB200_LIB=gpurun_exp/laps/libfiasco_b200.so timeout 300 python tools/gpu_check.py big > $OUT/r02b_single_laps_auto.txt 2>&1
tail -3 $OUT/r02b_single_laps_auto.txt
for c in 1 2; do
  FB200_CLUSTER=$c timeout 300 python tools/tiles_quick.py c4 3 > $OUT/r02b_c4_c$c.txt 2>&1; tail -1 $OUT/r02b_c4_c$c.txt
  FB200_CLUSTER=$c timeout 300 python tools/tiles_quick.py c3 3 > $OUT/r02b_c3_c$c.txt 2>&1; tail -1 $OUT/r02b_c3_c$c.txt
done
for c in 1 2 4 8; do
  FB200_CLUSTER=$c timeout 300 python tools/tiles_quick.py c2t 3 > $OUT/r02b_c2t_c$c.txt 2>&1; tail -1 $OUT/r02b_c2t_c$c.txt
done
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02b_pytest.txt 2>&1; tail -5 $OUT/r02b_pytest.txt

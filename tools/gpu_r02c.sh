#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r02c}
timeout 300 python tools/gpu_check.py big > $OUT/${TAG}_single_auto.txt 2>&1; grep g1024 $OUT/${TAG}_single_auto.txt
FB200_LIB=gpurun_exp/laps/libfiasco_b200.so timeout 300 python tools/gpu_check.py big > $OUT/${TAG}_single_laps.txt 2>&1
grep -A1 g1024 $OUT/${TAG}_single_laps.txt
for c in 1 2; do
  FB200_CLUSTER=$c timeout 300 python tools/tiles_quick.py c4 3 > $OUT/${TAG}_c4_c$c.txt 2>&1; tail -1 $OUT/${TAG}_c4_c$c.txt
done
FB200_CLUSTER=4 timeout 300 python tools/tiles_quick.py c2t 3 > $OUT/${TAG}_c2t_c4.txt 2>&1; tail -1 $OUT/${TAG}_c2t_c4.txt

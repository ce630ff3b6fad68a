#!/bin/bash
# experiment: thread-block shapes of the tile kernel (runs on the GPU box)
for cfg in ${CFGS:-"128 3 4" "128 2 4"}; do
  set -- $cfg
  FB200_NT=$1 FB200_BIG=$2 timeout 600 python tools/sweep_nt.py $3 74 2>&1 | tail -3
done

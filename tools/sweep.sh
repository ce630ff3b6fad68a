#!/bin/bash
# experiment: variants of the tile kernel (runs on the GPU box); LIBS = list of gpurun_exp/*.so,
# SHAPES = list of "NT:per_sm", BIGS = list of FB200_BIG values (bit 0: Gram rows global, bit 1: snapshots global)
for lib in ${LIBS:-product}; do
  if [ "$lib" = product ]; then unset FB200_LIB; else export FB200_LIB=gpurun_exp/$lib.so; fi
  for shape in ${SHAPES:-128:4}; do
    for big in ${BIGS:-default}; do
      if [ "$big" = default ]; then unset FB200_BIG; else export FB200_BIG=$big; fi
      FB200_NT=${shape%%:*} timeout 600 python tools/sweep_nt.py ${shape##*:} ${DISTINCT:-74} 2>&1 | tail -3 | grep -v "laps: ctrl=0.0"
    done
  done
done

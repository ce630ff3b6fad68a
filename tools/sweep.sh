#!/bin/bash
# experiment: variants of the tile kernel (runs on the GPU box); LIBS = list of gpurun_exp/*.so,
# SHAPES = list of "NT:per_sm", CARVES = list of shared-memory carve-out percentages
for lib in ${LIBS:-product}; do
  if [ "$lib" = product ]; then unset FB200_LIB; else export FB200_LIB=gpurun_exp/$lib.so; fi
  for shape in ${SHAPES:-128:4}; do
    for cv in ${CARVES:-default}; do
      if [ "$cv" = default ]; then unset FB200_CARVE; else export FB200_CARVE=$cv; fi
      echo "carve=$cv"
      FB200_NT=${shape%%:*} timeout 600 python tools/sweep_nt.py ${shape##*:} ${DISTINCT:-74} 2>&1 | tail -2 | grep -v "laps: ctrl=0.0"
    done
  done
done

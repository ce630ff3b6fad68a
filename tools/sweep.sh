#!/bin/bash
# experiment: variants of the tile kernel (runs on the GPU box); LIBS = list of gpurun_exp/*.so
for lib in ${LIBS:-product}; do
  if [ "$lib" = product ]; then unset FB200_LIB; else export FB200_LIB=gpurun_exp/$lib.so; fi
  FB200_NT=${NT:-128} timeout 600 python tools/sweep_nt.py ${PER_SM:-4} ${DISTINCT:-74} 2>&1 | tail -2
done

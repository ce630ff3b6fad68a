#!/bin/bash
# experiment helper: build the CUDA library of a git revision (default HEAD) into gpurun_exp/prev.so
# for A/B runs against the working tree (tools/sweep.sh LIBS="product prev")
set -e
rev=${1:-HEAD}
d=$(mktemp -d)
mkdir -p $d/fiasco_b200/csrc $d/include gpurun_exp
for f in fiasco_b200/csrc/tile_kernel.cu fiasco_b200/csrc/tile_kernel.cuh fiasco_b200/csrc/ffi.cu; do git show $rev:$f > $d/$f; done
cp include/*.h $d/include/
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -I$d/include -I$d/fiasco_b200/csrc"
$NV -c $d/fiasco_b200/csrc/tile_kernel.cu -o $d/tk.o
$NV -c $d/fiasco_b200/csrc/ffi.cu -o $d/ffi.o
extra=""
if git show $rev:fiasco_b200/csrc/motion_kernel.cu > $d/fiasco_b200/csrc/motion_kernel.cu 2>/dev/null; then
   $NV -c $d/fiasco_b200/csrc/motion_kernel.cu -o $d/mk.o && extra=$d/mk.o
fi
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_exp/prev.so $d/tk.o $d/ffi.o $extra -cudart static
rm -rf $d
echo "gpurun_exp/prev.so = $rev"

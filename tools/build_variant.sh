#!/bin/bash
# experiment helper: build the CUDA library with extra -D flags into gpurun_exp/<name>.so
# usage: tools/build_variant.sh name [-DFLAG ...]
set -e
name=$1; shift
mkdir -p gpurun_exp/obj_$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude -Ifiasco_b200/csrc"
$NV "$@" -Xptxas -v -c fiasco_b200/csrc/tile_kernel.cu -o gpurun_exp/obj_$name/tile_kernel.o 2> gpurun_exp/obj_$name/ptxas.log
$NV "$@" -c fiasco_b200/csrc/ffi.cu -o gpurun_exp/obj_$name/ffi.o
$NV "$@" -c fiasco_b200/csrc/motion_kernel.cu -o gpurun_exp/obj_$name/motion_kernel.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o gpurun_exp/$name.so gpurun_exp/obj_$name/tile_kernel.o gpurun_exp/obj_$name/ffi.o gpurun_exp/obj_$name/motion_kernel.o -cudart static
grep -E "spill" gpurun_exp/obj_$name/ptxas.log | sort | uniq -c | head -3
echo "$name: $(cuobjdump -sass gpurun_exp/obj_$name/tile_kernel.o | grep -c '^\s*/\*[0-9a-f]\{4,5\}\*/') sass"

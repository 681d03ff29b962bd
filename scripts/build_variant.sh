#!/bin/bash
# developer tool: build a variant of the library with extra nvcc flags -> curvedspacesim_b200/libvariant_<name>.so
# usage: scripts/build_variant.sh <name> [nvcc flags...]   ; run with CSS_LIB_PATH=<that file>
set -e
NAME=$1; shift
D=curvedspacesim_b200/csrc
OUT=/tmp/variant_$NAME; mkdir -p $OUT
COMMON="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
nvcc $COMMON -fmad=false "$@" -c $D/exact_kernels.cu -o $OUT/exact.o &
nvcc $COMMON "$@" -c $D/geodesic_kernel.cu -o $OUT/geo.o &
nvcc $COMMON "$@" -c $D/patch_kernel.cu -o $OUT/patch.o &
nvcc $COMMON "$@" -c $D/stencil_kernel.cu -o $OUT/stencil.o &
nvcc $COMMON "$@" -c $D/window_kernel.cu -o $OUT/win.o &
nvcc $COMMON "$@" -c $D/window_half_kernel.cu -o $OUT/winhalf.o &
nvcc $COMMON "$@" -c $D/css_api.cu -o $OUT/api.o &
nvcc $COMMON "$@" -c $D/microbench.cu -o $OUT/micro.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o curvedspacesim_b200/libvariant_$NAME.so $OUT/*.o -lnccl
echo built curvedspacesim_b200/libvariant_$NAME.so

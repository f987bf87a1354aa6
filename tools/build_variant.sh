#!/bin/bash
# builds lib/variants/libmvmc_<name>.so with extra -D flags for als.cu (A/B experiments; select with MVMC_LIBRARY=...)
set -e
name=$1; shift
D=multiview_motion_capture_b200
mkdir -p $D/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -I$D/csrc "$@" -c $D/csrc/als.cu -o $D/lib/variants/als_$name.o 2>&1 | grep -E "error" || true
nvcc -shared -o $D/lib/variants/libmvmc_$name.so $D/lib/affinity.o $D/lib/variants/als_$name.o $D/lib/assign.o $D/lib/ik.o $D/lib/pipeline.o -gencode arch=compute_100a,code=sm_100a
echo built $D/lib/variants/libmvmc_$name.so

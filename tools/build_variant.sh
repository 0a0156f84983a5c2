#!/bin/bash
# Development aid: build librt3d_<name>.so with extra -D flags for frame_batch.cu (kernel tuning experiments; select it
# with RT3D_LIB_PATH=<path>).  The shipped library is always rt-3d-reconstruction_b200/librt3d.so from the Makefile.
set -e
name=$1; shift
cd "$(dirname "$0")/../rt-3d-reconstruction_b200/csrc"
make -s -j8 > /dev/null
mkdir -p build/var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c frame_batch.cu -o build/var/frame_batch_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../librt3d_$name.so build/common.o build/k0_mask.o build/k1_backproject.o build/radix_sort.o build/k2_voxel.o build/k3_sor.o build/k4_fuse_subtract.o build/var/frame_batch_$name.o
cuobjdump --dump-resource-usage build/var/frame_batch_$name.o 2>/dev/null | grep -A1 "fb_project_tma_kernelILi2" | grep -oE "REG:[0-9]+" | head -1

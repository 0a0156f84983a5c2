#!/bin/bash
# Development aid: build librt3d_<name>.so with extra -D flags applied to EVERY source file (kernel tuning experiments; select it
# with RT3D_LIB_PATH=<path>).  The shipped library is always rt-3d-reconstruction_b200/librt3d.so from the Makefile.
set -e
name=$1; shift
cd "$(dirname "$0")/../rt-3d-reconstruction_b200/csrc"
d=build/var_$name; mkdir -p $d
for f in common k0_mask k1_backproject radix_sort k2_voxel k3_sor k4_fuse_subtract frame_batch; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c $f.cu -o $d/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../librt3d_$name.so $d/*.o

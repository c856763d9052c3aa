#!/bin/bash
# Build the C-ABI library for sm_100a in-tree (travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
PKG=deep-active-inference-mc_b200
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared \
     -o $PKG/libdai_b200.so $PKG/csrc/dai_api.cu $PKG/csrc/dai_simt.cu $PKG/csrc/dai_tc.cu $PKG/csrc/dai_frames.cu $PKG/csrc/dai_planner.cu \
     "$@"
echo "built $PKG/libdai_b200.so"

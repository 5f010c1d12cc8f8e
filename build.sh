#!/bin/bash
# Builds libscribe_b200.so (sm_100a only) in-tree; the .so travels to the GPU box with gpurun.
set -e
cd "$(dirname "$0")"
mkdir -p scribe_py_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" \
     -o scribe_py_b200/lib/libscribe_b200.so scribe_py_b200/csrc/capi.cu

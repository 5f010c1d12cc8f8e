#!/bin/bash
# tools/build_variant.sh NAME [nvcc flags...]  ->  tunelibs/libscribe_b200_NAME.so (a tuning build; select it with SCRIBE_B200_LIB)
set -e
cd "$(dirname "$0")/.."
name="$1"; shift
mkdir -p tunelibs
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" \
     -o "tunelibs/libscribe_b200_${name}.so" scribe_py_b200/csrc/capi.cu

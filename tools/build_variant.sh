#!/bin/bash
# Developer aid: link an experimental build of ONE translation unit into a copy of the library (same C ABI), e.g.
#   tools/build_variant.sh scalar npde_sep_m5.cu -DBODE_PAIR_SCALAR      ->  tools/ubench/libbode_scalar.so
# and load it with BODE_LIB_PATH=tools/ubench/libbode_scalar.so (bayesian-ode_b200/_lib.py).
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/../bayesian-ode_b200/csrc"
obj=/tmp/var_${name}_${src%.cu}.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c $src -o $obj 2> /tmp/var_${name}.ptxas.log
objs=$(ls *.o | grep -v "^${src%.cu}.o$")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/ubench/libbode_${name}.so $objs $obj -lcudart
echo built tools/ubench/libbode_${name}.so

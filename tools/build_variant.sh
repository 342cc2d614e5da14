#!/bin/bash
# A/B build of the library: tools/build_variant.sh NAME "-DMPRES_ALIGN_BLOCKS=4 ..."  ->  mpres-blas_b200/libmpres_b200_NAME.so
# (select it with MPRES_B200_LIB=$PWD/mpres-blas_b200/libmpres_b200_NAME.so; the other translation units come from the regular build)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -c -Xcompiler -fPIC $@ mpres-blas_b200/csrc/mpres_b200.cu -o mpres-blas_b200/build/variant_$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a mpres-blas_b200/build/variant_$name.o mpres-blas_b200/build/mpres_ops.o -o mpres-blas_b200/libmpres_b200_$name.so
echo "built $name"

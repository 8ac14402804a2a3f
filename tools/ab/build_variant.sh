#!/bin/bash
# Build an A/B variant of the library: tools/ab/build_variant.sh NAME FILE.cu "-DFLAG=..." -> tools/ab/lib_NAME.so
# (the other objects are taken from the regular build; select at run time with PLT_B200_LIB)
set -e
name=$1; src=$2; flags=$3
cd "$(dirname "$0")/../../polatory_b200/csrc"
mkdir -p build/ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function \
  --expt-relaxed-constexpr $flags -c $src -o build/ab/${name}.o
objs=$(ls build/*.o | grep -v "build/${src%.cu}.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/ab/lib_${name}.so $objs build/ab/${name}.o -lcudart -lpthread
echo built tools/ab/lib_${name}.so

#!/bin/bash
# builds variants/libiqw_<name>.so: the library with ONE source recompiled with extra flags (kernel experiments;
# select it with IQW_B200_LIB=variants/libiqw_<name>.so).  usage: tools/build_variant.sh name source.cu "-DFLAG=1 ..."
set -e
cd "$(dirname "$0")/../iqwaveform_b200/csrc"
name=$1; src=$2; flags=$3
mkdir -p ../../variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
obj=/tmp/iqw_variant_${name}.o
$NVCC -O3 -std=c++17 -lineinfo -DIQW_PACKED_F32X2 $ARCH -Xcompiler -fPIC,-O3,-Wall -Xptxas -v $flags -c -o $obj $src 2> /tmp/iqw_variant_${name}.log || (cat /tmp/iqw_variant_${name}.log; exit 1)
others=$(ls *.o | grep -v "^${src%.cu}.o$")
$NVCC $ARCH -shared --cudart shared -o ../../variants/libiqw_${name}.so $obj $others
echo built variants/libiqw_${name}.so

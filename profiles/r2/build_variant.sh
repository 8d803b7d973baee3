#!/bin/bash
# build a library variant for A/B runs: profiles/r2/build_variant.sh <name> <source.cu> [-D...]
#   -> blueice_b200/build/variants/lib_<name>.so  (select with BLUEICE_B200_LIB=<path>)
cd "$(dirname "$0")/../.." || exit 1
name=$1; src=$2; shift 2
mkdir -p blueice_b200/build/variants
obj=blueice_b200/build/variants/${name}_$(basename ${src} .cu).o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 "$@" \
    -c blueice_b200/csrc/${src} -o ${obj} || exit 1
others=$(ls blueice_b200/build/*.o | grep -v "/$(basename ${src} .cu).o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o blueice_b200/build/variants/lib_${name}.so ${obj} ${others} || exit 1
echo blueice_b200/build/variants/lib_${name}.so

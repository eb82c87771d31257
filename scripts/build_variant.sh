#!/bin/bash
# Experiment helper: build variants/libulcb200_<name>.so with extra nvcc flags for ONE source file (the other objects are the
# current build's).  usage: scripts/build_variant.sh <name> <file.cu> <extra flags...>; run with ULCB_LIB=variants/libulcb200_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v -I include "$@" \
     -c ulc_codec_b200/csrc/$src -o variants/${src%.cu}_$name.o 2>&1 | grep -A2 "transform_kernelILb1ELb0\|$VARIANT_GREP" | grep -E "spill|Used" | head -4
objs=$(ls ulc_codec_b200/build/*.o | grep -v "/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libulcb200_$name.so $objs variants/${src%.cu}_$name.o
echo built variants/libulcb200_$name.so

#!/bin/bash
# Build a variant of the library for kernel experiments: build/libspgg_<tag>.so with extra -D flags.
# usage: bash scripts/build_variant.sh <tag> [-DSPGG_X_... ...]     (select it with SPGG_B200_LIB=build/libspgg_<tag>.so)
TAG=$1; shift
PKG=$(ls -d neighbor*_b200)
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -w "$@" \
  -o build/libspgg_${TAG}.so $PKG/csrc/spgg_capi.cu $PKG/csrc/spgg_inst_general.cu $PKG/csrc/spgg_inst_lean.cu && echo "built build/libspgg_${TAG}.so"

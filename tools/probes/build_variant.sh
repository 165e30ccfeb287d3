#!/bin/bash
# usage: tools/probes/build_variant.sh <tag> [-DFLAG=V ...]   ->  tools/probes/lib_<tag>.so (developer experiments)
TAG=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared "$@" \
  -o tools/probes/lib_$TAG.so footsies_gym_b200/csrc/footsies_kernels.cu

#!/bin/bash
# A/B builds of the CUDA library with extra -D switches: tools/ab_build.sh <name> [-DFLAG=..]...  ->  build_ab/libavg_<name>.so
# (git-ignored; travels to the GPU box; select with AVG_B200_LIB=build_ab/libavg_<name>.so)
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build_ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -Xcompiler -fPIC -shared -std=c++17 "$@" \
  -o build_ab/libavg_$name.so assistive_vr_gym_b200/csrc/avg_kernels.cu assistive_vr_gym_b200/csrc/avg_capi.cu

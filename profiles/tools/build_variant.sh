#!/bin/bash
# Builds an A/B variant of libdtof_b200.so into exp_build/<name>.so: build_variant.sh <name> [extra nvcc flags...]
# (bench.py / the tests pick a variant with DTOF_LIB=$PWD/exp_build/<name>.so)
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
mkdir -p exp_build
cd mitsuba3dopplertof_b200/csrc
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-O2,-Wall -Xptxas -v \
  "$@" -shared dtof_api.cu dtof_bvh.cpp -o ../../exp_build/$name.so -lcudart -ldl 2> ../../exp_build/$name.ptxas.txt
grep -A2 "render_kernelILi1ELb0ELb0ELi0ELb0" ../../exp_build/$name.ptxas.txt | tail -2

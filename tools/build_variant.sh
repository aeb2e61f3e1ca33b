#!/bin/bash
# usage: tools/build_variant.sh <tag> [extra nvcc flags...]   -> bullet_envs_b200/csrc/variants/libsnake_b200_<tag>.so
# compile-time variants of the library for same-box A/B runs (load with SNK_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/../bullet_envs_b200/csrc"
mkdir -p variants
tag=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC,-pthread "$@" \
  -o variants/libsnake_b200_$tag.so snake_exact.cu snake_pgs.cu snake_gae.cu snake_abi.cu
echo built variants/libsnake_b200_$tag.so

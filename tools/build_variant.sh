#!/bin/bash
# usage: tools/build_variant.sh NAME [nvcc flags...]   ->  build/libnsr_NAME.so  (A/B runs: NSR_LIB_PATH=build/libnsr_NAME.so)
set -e
cd "$(dirname "$0")/../neural-sim-nerf_b200"
name=$1; shift
mkdir -p ../build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" -o ../build/libnsr_$name.so \
  csrc/api.cu csrc/ray_stage.cu csrc/image_stage.cu csrc/train_stage.cu csrc/mlp_forward.cu csrc/mlp_backward.cu csrc/wgrad.cu csrc/refine.cu
echo build/libnsr_$name.so

#!/bin/bash
# Builds libefusion.so (the product: sm_100a CUDA kernels + C ABI) in-tree. No reference or oracle code is linked.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRC="elasticfusion_b200/csrc/ef_api.cu elasticfusion_b200/csrc/ef_track.cu elasticfusion_b200/csrc/ef_map.cu elasticfusion_b200/csrc/ef_preprocess.cu"
OUT=elasticfusion_b200/libefusion.so
$NVCC -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false \
  -Xcompiler -fPIC,-O2,-Wall -ccbin /usr/bin/g++ -shared -o $OUT $SRC -lcudart "$@"
echo "built $OUT"

#!/bin/bash
# A/B builds: the same library with extra -D switches, next to the product one (selected at run time with
# GH_LIB_PATH=genhancer_b200/libgenhancer_b200_ab.so).   usage: tools/build_ab.sh -DGH_TEMPTY_RELEASE [...]
set -e
cd "$(dirname "$0")/../genhancer_b200/csrc"
mkdir -p build_ab
for f in api gemm_sm100 conv_sm100 patch_embed_sm100 attn_sm100 elementwise lora_fused norm vision optim; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function \
    --expt-relaxed-constexpr "$@" -c $f.cu -o build_ab/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libgenhancer_b200_ab.so build_ab/*.o
echo "built ../libgenhancer_b200_ab.so with $*"

#!/bin/bash
# build kernel variants into scratch/variants/<name>.so :  name:"-DTIC_CTAS=8 -DTIC_PRIV=8" ...
set -e
mkdir -p scratch/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v $flags \
     -o scratch/variants/$name.so tinyimgcodec_b200/csrc/tic_encode.cu 2> scratch/variants/$name.log &
done
wait
for spec in "$@"; do name=${spec%%:*}; echo "$name: $(grep -A2 'encode_tiles_kernelILb0' scratch/variants/$name.log | grep -E 'Used|spill' | tr '\n' ' ')"; done

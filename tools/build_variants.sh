#!/bin/bash
# build kernel variants into scratch/variants/<name>.so :  name:"-DTIC_GROUPS=5 -DTIC_PREFETCH=0" ...
# (the decode side is linked from the object of the regular build: run `python -m tinyimgcodec_b200.build` first)
set -e
mkdir -p scratch/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v $flags \
     -c -o scratch/variants/$name.o tinyimgcodec_b200/csrc/tic_encode.cu 2> scratch/variants/$name.log &&
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o scratch/variants/$name.so scratch/variants/$name.o \
     tinyimgcodec_b200/csrc/_obj/tic_decode.o ) &
done
wait
for spec in "$@"; do name=${spec%%:*}; echo "$name: $(grep -A2 'encode_tiles_kernelILi0' scratch/variants/$name.log | grep -E 'Used|spill' | tr '\n' ' ')"; done

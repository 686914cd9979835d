#!/bin/bash
# full ncu captures of the three big decode kernels on 1024 images (each after the plain command exited 0)
set -u
TAG=${1:-dec}
mkdir -p gpurun_out
DPROF="python tools/decode_bench.py --images 1024 --steps 2"
timeout 120 $DPROF > gpurun_out/dec_plain_${TAG}.json 2> gpurun_out/dec_plain_${TAG}.err; echo "decode plain rc=$?"; cut -c1-400 gpurun_out/dec_plain_${TAG}.json
for K in dec_sync_kernel dec_write_kernel dec_idct_fast_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 \
      -o gpurun_out/prof_${TAG}_$K -f $DPROF > gpurun_out/ncu_${TAG}_$K.log 2>&1
  echo "full capture $K rc=$?"
done
cp tinyimgcodec_b200/libtinyimgcodec_cuda.so gpurun_out/lib_${TAG}.so

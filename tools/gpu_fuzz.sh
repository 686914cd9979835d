#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -q --timeout=120 -k "never_crash" 2>&1 | tail -15
PROF="python tools/decode_bench.py --images 1024 --steps 2"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dec_idct_fast_kernel -s 0 -c 1 \
      -o gpurun_out/prof_r1l_dec_idct_fast_kernel -f $PROF > gpurun_out/dec_ncu_full_r1l_dec_idct_fast_kernel.log 2>&1
echo "fast idct capture rc=$?"

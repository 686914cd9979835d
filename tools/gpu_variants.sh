#!/bin/bash
# One GPU visit: [the GPU test suite on the in-tree library,] the short bench once per library in scratch/variants,
# and (TAG given) a full ncu capture of the encode kernel of the in-tree library.
# usage: tools/gpu_variants.sh [test|notest] [TAG]
set -u
mkdir -p gpurun_out
if [ "${1:-test}" = "test" ]; then timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -6; fi
tools/bench_variants.sh
TAG=${2:-}
if [ -n "$TAG" ]; then
  PROF="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-decode"
  $PROF > gpurun_out/plain_${TAG}.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_tiles_kernel -s 3 -c 1 \
      -o gpurun_out/prof_${TAG} -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
  echo "full capture rc=$?"
  cp tinyimgcodec_b200/libtinyimgcodec_cuda.so gpurun_out/lib_${TAG}.so
  mkdir -p gpurun_out/src_${TAG} && cp tinyimgcodec_b200/csrc/*.cu* gpurun_out/src_${TAG}/
fi

#!/bin/bash
# decode-side GPU iteration: debug cases, decode parity tests, then the whole GPU suite
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_debug_decode.py > gpurun_out/dec_debug.log 2>&1; echo "debug rc=$?"; tail -20 gpurun_out/dec_debug.log
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -x -q --timeout=300 2>&1 | tail -25

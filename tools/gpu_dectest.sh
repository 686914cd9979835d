#!/bin/bash
# flakiness hunt: the first three GPU test files (suite order) N times, stop at the first failing run and keep its log
mkdir -p gpurun_out
N=${1:-6}
for i in $(seq 1 $N); do
  timeout 300 python -m pytest tests/test_gpu_cli.py tests/test_gpu_cvariant.py tests/test_gpu_decode.py -m gpu -x -q --timeout=120 > gpurun_out/flake_$i.log 2>&1
  rc=$?; echo "run $i rc=$rc $(tail -1 gpurun_out/flake_$i.log)"
  if [ $rc -ne 0 ]; then grep -n "^E \|^FAILED\|Error" gpurun_out/flake_$i.log | head -30; break; fi
done

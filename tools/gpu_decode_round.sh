#!/bin/bash
# decode-side GPU iteration: debug cases, decode parity tests, short bench with the decode leg
set -u
mkdir -p gpurun_out
timeout 300 python tools/gpu_debug_decode.py > gpurun_out/dec_debug.log 2>&1; echo "debug rc=$?"; grep -c OK gpurun_out/dec_debug.log; grep -v " OK " gpurun_out/dec_debug.log | tail; grep flat2048 gpurun_out/dec_debug.log
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -x -q --timeout=300 2>&1 | tail -25
timeout 300 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/dec_quick.json 2> gpurun_out/dec_quick.err || tail -5 gpurun_out/dec_quick.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/dec_quick.json"))
print("ENCODE step_ms=%.3f" % d["ms_per_step"], "DECODE", json.dumps(d["decode"]))
PY

#!/bin/bash
# decode-side GPU iteration: debug cases (short timeout: a hang must not eat the GPU budget — the visit stops there),
# decode parity tests, short bench with the decode leg
set -u
mkdir -p gpurun_out
timeout 90 python tools/gpu_debug_decode.py > gpurun_out/dec_debug.log 2>&1; rc=$?; echo "debug rc=$rc"; grep -c OK gpurun_out/dec_debug.log; grep -v " OK " gpurun_out/dec_debug.log | tail; grep flat2048 gpurun_out/dec_debug.log
if [ $rc -ne 0 ]; then echo "debug cases failed or hung: stopping"; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_decode.py -m gpu -x -q --timeout=120 2>&1 | tail -25
timeout 200 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/dec_quick.json 2> gpurun_out/dec_quick.err || tail -5 gpurun_out/dec_quick.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/dec_quick.json"))
print("ENCODE step_ms=%.3f" % d["ms_per_step"], "DECODE", json.dumps(d["decode"]))
PY

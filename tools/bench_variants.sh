#!/bin/bash
# run the short bench once per variant library in scratch/variants
for f in scratch/variants/*.so; do
  n=$(basename $f .so)
  TIC_LIB_PATH=$PWD/$f timeout 300 python bench.py --no-cpu --no-e2e --no-decode --steps 10 --warmup 3 > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/var_{n}.json"))
    print("VARIANT %-24s step_ms=%.3f kernel_ms=%.3f compact_ms=%.3f parity=%s" % (n, d["ms_per_step"], d["roofline"]["launch_ms"], d["roofline"]["other_kernels_ms"]["compact_kernel"], d["parity"]))
except Exception as e:
    print("VARIANT", n, "failed", e)
PY
done

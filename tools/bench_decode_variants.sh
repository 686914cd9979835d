#!/bin/bash
# run the short bench (with the decode leg) once per variant library in scratch/variants
for f in scratch/variants/*.so; do
  n=$(basename $f .so)
  TIC_LIB_PATH=$PWD/$f timeout 200 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/var_{n}.json")); dec=d["decode"]
    print("VARIANT %-10s enc=%.3f dec=%.3f %s rounds=%s parity=%s" % (n, d["ms_per_step"], dec["ms_per_step"], {k: round(v, 2) for k, v in dec["phases_ms"].items()}, dec["sync_rounds"], dec["parity_vs_oracle"]))
except Exception as e:
    print("VARIANT", n, "failed", e)
PY
done

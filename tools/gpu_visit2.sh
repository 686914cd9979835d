#!/bin/bash
# one visit: encode variants (scratch/variants), auto-table and decode variants (scratch/variants2)
set -u
mkdir -p gpurun_out
tools/bench_variants.sh
for v in base statsdirect; do
  TIC_LIB_PATH=$PWD/scratch/variants2/$v.so python tools/auto_bench.py --images 1024 --steps 4 > gpurun_out/auto_$v.json 2> gpurun_out/auto_$v.err
  python - $v <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/auto_{sys.argv[1]}.json"))
print("AUTO", sys.argv[1], [round(x,3) for x in d["auto"]["ms"]], "bytes", d["auto"]["bytes"])
PY
done
for v in base dec64; do
  TIC_LIB_PATH=$PWD/scratch/variants2/$v.so python tools/decode_bench.py --images 1024 --steps 4 > gpurun_out/dec_$v.json 2> gpurun_out/dec_$v.err
  cut -c1-420 gpurun_out/dec_$v.json
done

#!/bin/bash
# quick GPU iteration: parity tests, short bench, optional full ncu capture (TAG given => capture)
set -u
TAG=${1:-}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q --timeout=120 2>&1 | tail -15
timeout 180 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/quick.json 2> gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/quick.json"))
print("BENCH value=%.0f Mpx/s step_ms=%.3f kernel_ms=%.3f compact_ms=%.3f frac=%.4f parity=%s clocks=%s exact=%s" % (d["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["roofline"]["other_kernels_ms"]["compact_kernel"], d["roofline"]["frac"], d["parity"], d["clocks"], d["config"]["exact_path"]))
PY
if [ -n "$TAG" ]; then
  PROF="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"compact_kernel|encode_tiles_kernel|finalize_kernel|scan_|symbol_stats|build_tables" -c 40 --csv \
      --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launch_${TAG}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_tiles_kernel -s 3 -c 1 \
      -o gpurun_out/prof_${TAG} -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
  echo "full capture rc=$?"
  cp tinyimgcodec_b200/libtinyimgcodec_cuda.so gpurun_out/lib_${TAG}.so
fi

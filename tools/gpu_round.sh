#!/bin/bash
# One GPU-box visit for the record: GPU test suite, full bench (both arms), launch lists, full ncu captures,
# the other BASELINE configurations.  Outputs in gpurun_out/.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_${TAG}.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
echo "reference arm rc=$?"; tail -c 600 gpurun_out/bench_ref_${TAG}.json
PROF="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
KERNELS='regex:compact_kernel|encode_tiles_kernel|finalize_kernel|scan_|dec_'
$PROF > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 120 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
if [ "${2:-}" = "encode" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_tiles_kernel -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_encode -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture (encode) rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compact_kernel -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_compact -f $PROF > gpurun_out/ncu_full2_${TAG}.log 2>&1
echo "full capture (compact) rc=$?"
fi
DPROF="python tools/decode_bench.py --images 1024 --steps 2"
$DPROF > gpurun_out/dec_plain_${TAG}.json 2> gpurun_out/dec_plain_${TAG}.err; echo "decode plain rc=$?"; cat gpurun_out/dec_plain_${TAG}.json
for K in dec_idct_kernel dec_write_kernel dec_sync_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 \
      -o gpurun_out/prof_${TAG}_$K -f $DPROF > gpurun_out/dec_ncu_full_${TAG}_$K.log 2>&1
  echo "full capture $K rc=$?"
done
timeout 900 python tools/bench_configs.py > gpurun_out/configs_${TAG}.jsonl 2> gpurun_out/configs_${TAG}.err; echo "configs rc=$?"
cut -c1-400 gpurun_out/configs_${TAG}.jsonl
cp tinyimgcodec_b200/libtinyimgcodec_cuda.so gpurun_out/lib_${TAG}.so

#!/bin/bash
# One GPU-box visit for the record: full bench (both arms), launch list, full ncu captures.  Outputs in gpurun_out/.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
echo "reference arm rc=$?"; tail -c 600 gpurun_out/bench_ref_${TAG}.json
PROF="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
KERNELS='regex:compact_kernel|encode_tiles_kernel|finalize_kernel|scan_'
$PROF > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 30 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_tiles_kernel -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_encode -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture (encode) rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compact_kernel -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_compact -f $PROF > gpurun_out/ncu_full2_${TAG}.log 2>&1
echo "full capture (compact) rc=$?"
cp tinyimgcodec_b200/libtinyimgcodec_cuda.so gpurun_out/lib_${TAG}.so

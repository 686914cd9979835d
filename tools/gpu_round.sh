#!/bin/bash
# One GPU-box visit: bench, launch list, full ncu capture of the encode kernel.  Outputs in gpurun_out/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json
PROF="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
$PROF > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'encode_tiles_kernel|finalize_kernel' -c 12 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
$PROF > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:encode_tiles_kernel -s 3 -c 1 \
    -o gpurun_out/prof_${TAG} -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_full_${TAG}.log

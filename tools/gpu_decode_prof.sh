#!/bin/bash
# decode-side profile: launch list + one full ncu capture of each decode kernel
set -u
TAG=${1:-r1k}
mkdir -p gpurun_out
PROF="python tools/decode_bench.py --images 1024 --steps 2"
$PROF > gpurun_out/dec_plain_${TAG}.json 2> gpurun_out/dec_plain_${TAG}.err; echo "plain rc=$?"; cat gpurun_out/dec_plain_${TAG}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dec_" -c 40 --csv \
    --log-file gpurun_out/dec_launches_${TAG}.csv $PROF > gpurun_out/dec_ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
for K in dec_idct_kernel dec_write_kernel dec_sync_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 \
      -o gpurun_out/prof_${TAG}_$K -f $PROF > gpurun_out/dec_ncu_full_${TAG}_$K.log 2>&1
  echo "full capture $K rc=$?"
done

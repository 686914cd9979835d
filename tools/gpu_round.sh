#!/bin/bash
# One GPU-box visit for the record: GPU test suite, full bench (both arms), launch list, full ncu captures of every
# kernel (each only after the same command has exited 0 without ncu), the other BASELINE configurations.
# Outputs in gpurun_out/.   usage: tools/gpu_round.sh <tag> [encode|all]
set -u
TAG=${1:-r2}
WHAT=${2:-all}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
echo "reference arm rc=$?"; tail -c 900 gpurun_out/bench_ref_${TAG}.json
PROF="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-decode"
KERNELS='regex:compact_kernel|encode_tiles_kernel|prep_uniform|scan_|symbol_stats|build_tables'
$PROF > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 160 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $PROF > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_tiles_kernel -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_encode -f $PROF > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture (encode<0>) rc=$?"
if [ "$WHAT" = "all" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compact_kernel -s 3 -c 1 \
    -o gpurun_out/prof_${TAG}_compact -f $PROF > gpurun_out/ncu_full2_${TAG}.log 2>&1
echo "full capture (compact) rc=$?"
APROF="python tools/auto_bench.py --images 1024 --steps 1"
$APROF > gpurun_out/auto_plain_${TAG}.json 2> gpurun_out/auto_plain_${TAG}.err; echo "auto plain rc=$?"; cut -c1-600 gpurun_out/auto_plain_${TAG}.json
for K in symbol_stats build_tables_kernel coeffs_kernel scan_chunks_kernel scan_apply_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 \
      -o gpurun_out/prof_${TAG}_$K -f $APROF > gpurun_out/auto_ncu_${TAG}_$K.log 2>&1
  echo "full capture $K rc=$?"
done
# encode_tiles_kernel<1> (auto) is the 1st and <2> (C variant) the 3rd..4th encode launch of that driver
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_tiles_kernel -s 0 -c 1 \
    -o gpurun_out/prof_${TAG}_encode_auto -f $APROF > gpurun_out/auto_ncu_${TAG}_enc1.log 2>&1
echo "full capture encode<1> rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_tiles_kernel -s 2 -c 1 \
    -o gpurun_out/prof_${TAG}_encode_cvar -f $APROF > gpurun_out/auto_ncu_${TAG}_enc2.log 2>&1
echo "full capture encode<2> rc=$?"
DPROF="python tools/decode_bench.py --images 1024 --steps 2"
$DPROF > gpurun_out/dec_plain_${TAG}.json 2> gpurun_out/dec_plain_${TAG}.err; echo "decode plain rc=$?"; cut -c1-500 gpurun_out/dec_plain_${TAG}.json
timeout 900 python tools/bench_configs.py > gpurun_out/configs_${TAG}.jsonl 2> gpurun_out/configs_${TAG}.err; echo "configs rc=$?"
cut -c1-300 gpurun_out/configs_${TAG}.jsonl
fi
cp tinyimgcodec_b200/libtinyimgcodec_cuda.so gpurun_out/lib_${TAG}.so
mkdir -p gpurun_out/src_${TAG} && cp tinyimgcodec_b200/csrc/*.cu* gpurun_out/src_${TAG}/

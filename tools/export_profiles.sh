#!/bin/bash
# gpurun_out/ of one record visit (tools/gpu_round.sh <tag> all) -> profiles/<tag>_*: bench lines, launch list, raw ncu
# exports per kernel, the per-line summary of the encode kernel, the kernel table, the SASS histogram.
# usage: tools/export_profiles.sh <tag>
set -eu
T=$1; G=gpurun_out; P=profiles
cp $G/bench_$T.json $P/${T}_bench.json
cp $G/bench_ref_$T.json $P/${T}_bench_reference_arm.json
cp $G/launches_$T.csv $P/${T}_launches.csv
cp $G/configs_$T.jsonl $P/${T}_other_configs.jsonl
cp $G/dec_plain_$T.json $P/${T}_decode_1024_images.json
cp $G/auto_plain_$T.json $P/${T}_auto_cvariant_1024_images.json
declare -A NAME=( [encode]=encode [compact]=compact [encode_auto]=encode_auto [encode_cvar]=encode_cvar
  [symbol_stats]=symbol_stats_kernel [build_tables_kernel]=build_tables_kernel [coeffs_kernel]=coeffs_kernel
  [scan_chunks_kernel]=scan_chunks_kernel [scan_apply_kernel]=scan_apply_kernel )
for k in "${!NAME[@]}"; do
  f=$G/prof_${T}_$k.ncu-rep
  [ -f $f ] && ncu -i $f --page raw --csv > $P/${T}_${NAME[$k]}_ncu_raw.csv 2>/dev/null
done
ncu -i $G/prof_${T}_encode.ncu-rep --page source --csv --print-source sass > $G/${T}_encode_sass.csv 2>/dev/null
{ python tools/ncu_src_summary.py $G/${T}_encode_sass.csv 25
  echo; echo "=== by source line / function (sources as built: gpurun_out/src_$T) ==="
  python tools/ncu_by_line.py $G/${T}_encode_sass.csv $G/lib_$T.so encode_tiles_kernelILi0ELi7 40; } > $P/${T}_encode_tiles_source_summary.txt
python tools/ncu_kernel_table.py $P/${T}_*_ncu_raw.csv > $P/${T}_kernel_table.md
python tools/sass_histogram.py $G/lib_$T.so > $P/${T}_sass_histogram.txt
ls -la $P | grep ${T}_

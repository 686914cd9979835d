#!/bin/bash
# per library in scratch/variants: the short bench, then the high-bitrate configurations (C5 q90/q80, S1)
set -u
mkdir -p gpurun_out
for f in scratch/variants/*.so; do
  n=$(basename $f .so)
  TIC_LIB_PATH=$PWD/$f timeout 200 python bench.py --no-cpu --no-e2e --no-decode --steps 10 --warmup 3 > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python -c "
import json; d=json.load(open('gpurun_out/var_$n.json')); print('VARIANT $n kernel_ms=%.3f step=%.3f parity=%s' % (d['roofline']['launch_ms'], d['ms_per_step'], d['parity']))"
  TIC_LIB_PATH=$PWD/$f timeout 300 python tools/bench_configs.py S1 C2 > gpurun_out/cfg_$n.jsonl 2> gpurun_out/cfg_$n.err
  python -c "
import json
for l in open('gpurun_out/cfg_$n.jsonl'):
    d=json.loads(l); print('   %-34s q%-3s ms %.3f kernel %.3f' % (d['config'][:34], d['quality'], d['ms_median'], d.get('encode_kernel_ms',0)))"
done

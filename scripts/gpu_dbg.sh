#!/bin/bash
# timing-only experiments (DV_TC_DEBUG produces wrong results on purpose): per-entry-point eager times
set -u
mkdir -p gpurun_out
for v in 0 1 2; do
  DV_TC_DEBUG=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dbg$v.json 2> gpurun_out/bench_dbg$v.err; echo "dbg $v exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_dbg$v.json'))
print('debug=$v ms', d['ms_per_step'], 'profiled', d.get('profiled_call_ms_per_step'), d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('us_per_launch'))
print('  share', d.get('kernel_share'))"
done

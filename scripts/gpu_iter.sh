#!/bin/bash
# quick iteration: conv + model tests, bench (no cpu baseline), optional ncu captures
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout=300 ${PYTEST_ARGS:-} > gpurun_out/iter_tests.log 2>&1; echo "tests exit $?"
grep -E "passed|failed|rel err|Error" gpurun_out/iter_tests.log | head -12
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench exit $?"; tail -2 gpurun_out/bench_iter.err
python -c "
import json; d=json.load(open('gpurun_out/bench_iter.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('share', d.get('kernel_share')); print('logdens', d.get('roofline_logdensity',{}).get('us'), d.get('roofline_logdensity',{}).get('frac'))"
for k in ${KERNELS:-}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_SKIP:-3} -c 1 -f -o gpurun_out/prof_$k python scripts/prof_step.py > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done

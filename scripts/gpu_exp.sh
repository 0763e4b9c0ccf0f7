#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout=300 -x > gpurun_out/iter_tests.log 2>&1; echo "tests exit $?"
grep -E "passed|failed|rel err|Error" gpurun_out/iter_tests.log | head -12
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json'))
print('$name value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('  share', d.get('kernel_share'))"
}
run base A=1
for v in ${VARIANTS:-}; do run "${v//=/_}" $v; done
timeout 300 python scripts/find_fills.py > gpurun_out/fills.log 2>&1; echo "fills exit $?"; head -45 gpurun_out/fills.log

#!/bin/bash
# linear tensor-core kernels: tests, then bench A/B against the FFMA path
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout=120 -k "linear" > gpurun_out/lin_tests.log 2>&1; echo "linear tests exit $?"
grep -E "passed|failed|rel err|Error" gpurun_out/lin_tests.log | head -20
timeout 600 python -m pytest tests -q -m gpu --timeout=300 -x > gpurun_out/iter_tests.log 2>&1; echo "all tests exit $?"
grep -E "passed|failed|rel err|Error" gpurun_out/iter_tests.log | head -12
for impl in tc ffma; do
  DV_LINEAR_IMPL=$impl timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lin_$impl.json 2> gpurun_out/bench_lin_$impl.err; echo "bench $impl exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_lin_$impl.json'))
print('$impl value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'parity', d.get('parity'))
print('share', d.get('kernel_share'))"
done
DV_LINEAR_IMPL=tc timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lin_c4.json 2>gpurun_out/bench_lin_c4.err; echo "c4 exit $?"
DV_LINEAR_IMPL=ffma timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lin_c4_ffma.json 2>gpurun_out/bench_lin_c4_ffma.err; echo "c4 ffma exit $?"
python -c "
import json
for f in ['gpurun_out/bench_lin_c4.json','gpurun_out/bench_lin_c4_ffma.json']:
    d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d.get('parity')); print(d.get('kernel_share'))"

#!/bin/bash
# tcgen05 bring-up: conv parity with the tensor-core path, then everything, then bench.
set -u
mkdir -p gpurun_out
echo "== conv tests (tc: down+up only)"
DV_TC_DISABLE=wgrad timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv" --timeout=120 > gpurun_out/tc_conv_nowg.log 2>&1; echo "exit $?" | tee -a gpurun_out/tc_conv_nowg.log
tail -12 gpurun_out/tc_conv_nowg.log | grep -E "passed|failed|rel err|Error" 
echo "== conv tests (tc: all)"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv" --timeout=120 > gpurun_out/tc_conv.log 2>&1; echo "exit $?" | tee -a gpurun_out/tc_conv.log
grep -E "passed|failed|rel err|Error" gpurun_out/tc_conv.log | head -20
echo "== all gpu tests (tc)"
timeout 900 python -m pytest tests -q -m gpu --timeout=300 > gpurun_out/tc_all.log 2>&1; echo "exit $?" | tee -a gpurun_out/tc_all.log
grep -E "passed|failed|rel err|Error" gpurun_out/tc_all.log | head -20
echo "== bench tc"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench exit $?"; tail -3 gpurun_out/bench_tc.err; cat gpurun_out/bench_tc.json

#!/bin/bash
# One gpurun call: parity tests, smoke, a short bench, an ncu launch list.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" 
timeout 1500 python -m pytest tests -q -m gpu -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps ${BENCH_STEPS:-20} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${RUN_NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
fi
if [ "${RUN_REF:-0}" = "1" ]; then
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
fi

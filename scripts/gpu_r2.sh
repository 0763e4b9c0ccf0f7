#!/bin/bash
# Round-2 GPU-box driver (one gpurun call): parity tests, smoke, bench (c2 + the other workloads), reference arms,
# ncu launch list.  Everything lands in gpurun_out/.  Env: TESTS=0 skips pytest, WORKLOADS="c3 c4 c5", NCU=0.
set -u
mkdir -p gpurun_out
# guard: the tcgen05 conv kernels of this configuration must terminate and be accurate before the long runs
if ! timeout 180 python scripts/sanity_ts.py; then echo "sanity_ts FAILED or hung: aborting this run"; exit 1; fi
timeout 240 python scripts/sanity_wg.py; rc=$?; if [ $rc -gt 1 ]; then echo "sanity_wg crashed or hung ($rc): aborting this run"; exit 1; fi
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${TESTS:-1}" = "1" ]; then
echo "== pytest gpu"
timeout 2400 python -m pytest tests -q -m gpu -x --timeout=900 --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
fi
echo "== bench c2 (full line)"
timeout 900 python bench.py --steps ${BENCH_STEPS:-100} --warmup 5 --detail > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -40 gpurun_out/bench.err; cat gpurun_out/bench.json
for w in ${WORKLOADS:-}; do
  echo "== bench $w"
  timeout 900 python bench.py --workload $w --steps 50 --warmup 5 --detail > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
  tail -25 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json
done
if [ "${RUN_REF:-1}" = "1" ]; then
echo "== reference arm (CPU)"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
fi
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-parity > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
fi

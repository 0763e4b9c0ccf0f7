set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout=900 --durations=10 -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python bench.py --steps 50 --warmup 5 --detail --no-cpu-baseline --no-eager-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -32 gpurun_out/bench.err; cat gpurun_out/bench.json

#!/bin/bash
# Quick GPU-box iteration: full parity suite, log-density phase timing, then bench.py (c2, no baselines) once per
# VARIANTS entry ("ENV=val" or "base").  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
# guard: the tcgen05 conv kernels of this configuration must terminate and be accurate before the long runs
if ! SANITY_TIME=1 timeout 180 python scripts/sanity_ts.py; then echo "sanity_ts FAILED or hung: aborting this run"; exit 1; fi
timeout 240 python scripts/sanity_wg.py; rc=$?; if [ $rc -gt 1 ]; then echo "sanity_wg crashed or hung ($rc): aborting this run"; exit 1; fi
if [ "${TESTS:-1}" = "1" ]; then
timeout 2400 python -m pytest tests -q -m gpu --timeout=900 --durations=8 -s ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|same branch|flipped" gpurun_out/pytest_gpu.log | tail -40
fi
echo "== btcvae timing"; python scripts/btcvae_timing.py 2>&1 | tail -4
for v in ${VARIANTS:-base}; do
  echo "== bench $v"
  if [ "$v" = "base" ]; then envs=""; else envs="$v"; fi
  env $envs timeout 600 python bench.py --steps 50 --warmup 5 --detail --no-cpu-baseline --no-eager-baseline ${BENCH_ARGS:-} > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench exit $?"
  grep -E "dv_conv|dv_btcvae|dv_linear" gpurun_out/bench_$v.err | head -24
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$v.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "launches/step", d.get("launches_per_step"))
print("logdensity", d.get("roofline_logdensity",{}).get("us"), "parity", json.dumps(d.get("parity")))
PY
done

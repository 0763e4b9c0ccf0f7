#!/bin/bash
# Round-2 multi-GPU driver (gpurun --gpus NG): data-parallel parity on real NCCL ranks (tests/ddp_worker.py: btcvae,
# FactorVAE, global-batch btcvae) and bench.py -- c2 weak scaling (ddp_parity block included) plus the two configs BASELINE
# assigns to 8 GPUs in strong scaling (c4: 512 -> 512/N per GPU, c5: 2048 -> 2048/N per GPU).  Outputs -> gpurun_out/.
# Env: NG, DDP=0 (skip the parity workers), JOBS="c2 weak;c5 strong", STEPS, EXTRA_ENV="VAR=val" (second c2 line under that environment).
set -u
NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
port=29500
if [ "${DDP:-1}" = "1" ]; then
for spec in "btcvae" "factor --img 3,64,64 --per 32" "btcvae --img 3,64,64 --z 64 --per 32 --global-btcvae"; do
  port=$((port+1))
  timeout 600 $TR --master-port $port tests/ddp_worker.py --loss $spec > gpurun_out/ddp_${NG}_$(echo $spec | tr ' ,-' '___').log 2>&1
  echo "ddp_worker [$spec] exit $?"; grep DDP_WORKER gpurun_out/ddp_${NG}_$(echo $spec | tr ' ,-' '___').log | cut -c1-400
done
fi
IFS=";" read -ra JOBLIST <<< "${JOBS:-c2 weak;c4 strong;c5 strong}"
for job in "${JOBLIST[@]}"; do
  set -- $job
  port=$((port+1))
  timeout 900 $TR --master-port $port bench.py --gpus $NG --workload $1 --scaling $2 --steps ${STEPS:-50} --warmup 5 \
    > gpurun_out/scale_${1}_${2}_n${NG}.json 2> gpurun_out/scale_${1}_${2}_n${NG}.err
  echo "bench $1 $2 N=$NG exit $?"; cut -c1-700 gpurun_out/scale_${1}_${2}_n${NG}.json
done
if [ -n "${EXTRA_ENV:-}" ]; then
  port=$((port+1))
  env $EXTRA_ENV timeout 900 $TR --master-port $port bench.py --gpus $NG --workload c2 --steps ${STEPS:-50} --warmup 5 \
    > gpurun_out/scale_c2_weak_n${NG}_extra.json 2> gpurun_out/scale_c2_weak_n${NG}_extra.err
  echo "bench c2 weak N=$NG [$EXTRA_ENV] exit $?"; cut -c1-400 gpurun_out/scale_c2_weak_n${NG}_extra.json
fi

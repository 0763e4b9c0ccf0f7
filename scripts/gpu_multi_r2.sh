#!/bin/bash
# Round-2 multi-GPU driver (gpurun --gpus NG): data-parallel parity on real NCCL ranks (tests/ddp_worker.py: btcvae,
# FactorVAE, global-batch btcvae) and bench.py -- c2 weak scaling (ddp_parity block included) plus the two configs BASELINE
# assigns to 8 GPUs in strong scaling (c4: 512 -> 512/N per GPU, c5: 2048 -> 2048/N per GPU).  Outputs -> gpurun_out/.
set -u
NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
port=29500
for spec in "btcvae" "factor --img 3,64,64 --per 32" "btcvae --img 3,64,64 --z 64 --per 32 --global-btcvae"; do
  port=$((port+1))
  timeout 600 $TR --master-port $port tests/ddp_worker.py --loss $spec > gpurun_out/ddp_${NG}_$(echo $spec | tr ' ,-' '___').log 2>&1
  echo "ddp_worker [$spec] exit $?"; grep DDP_WORKER gpurun_out/ddp_${NG}_$(echo $spec | tr ' ,-' '___').log | cut -c1-400
done
for job in "c2 weak" "c4 strong" "c5 strong"; do
  set -- $job
  port=$((port+1))
  timeout 900 $TR --master-port $port bench.py --gpus $NG --workload $1 --scaling $2 --steps ${STEPS:-50} --warmup 5 \
    > gpurun_out/scale_${1}_${2}_n${NG}.json 2> gpurun_out/scale_${1}_${2}_n${NG}.err
  echo "bench $1 $2 N=$NG exit $?"; cut -c1-700 gpurun_out/scale_${1}_${2}_n${NG}.json
done

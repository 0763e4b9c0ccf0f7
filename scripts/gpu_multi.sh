#!/bin/bash
# multi-GPU: DDP semantics check + weak-scaling bench at N = $NG
set -u
NG=${NG:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 scripts/ddp_check.py > gpurun_out/ddp_check_$NG.log 2>&1; echo "ddp_check exit $?"; tail -6 gpurun_out/ddp_check_$NG.log
for n in ${BENCH_NS:-$NG}; do
  if [ "$n" = "1" ]; then timeout 600 python bench.py --gpus 1 --steps 30 --warmup 6 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 30 --warmup 6 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  echo "bench N=$n exit $?"; tail -2 gpurun_out/scale_$n.err | cut -c1-300
  python -c "
import json; d=json.loads(open('gpurun_out/scale_$n.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'graph', d.get('cuda_graph'))"
done

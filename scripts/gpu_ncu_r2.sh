#!/bin/bash
# Round-2 ncu --set full captures (one launch each) from scripts/prof_step.py (eager steps, no graph).
# KERNELS="name:regex:skip ..." overrides the list.
set -u
mkdir -p gpurun_out
cap() {  # name regex skip
  DISVAE_CUDA_GRAPH=0 STEPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f \
    -o gpurun_out/r02_$1 python scripts/prof_step.py > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
}
for k in ${KERNELS:-img_down_fwd:img_down_kernel:0 img_down_masked:img_down_kernel:1 img_up:img_up_kernel:0 img_wgrad:img_wgrad_kernel:0 \
         btcvae_fwd4:btcvae_fwd4_kernel:0 wgrad32_h16:conv_wgrad32_ts_kernel:0 down32_h16:conv_down32_ts_kernel:0 up_halo32_h16:conv_up_halo_ts_kernel:2}; do
  IFS=: read name regex skip <<< "$k"
  cap $name $regex $skip
done
for f in gpurun_out/r02_*.ncu-rep; do
  n=$(basename $f .ncu-rep)
  python scripts/ncu_summary.py $f "$n (ncu --set full, --clock-control none)" > gpurun_out/$n.md 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep | head -20

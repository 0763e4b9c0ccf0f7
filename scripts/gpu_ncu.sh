#!/bin/bash
# ncu --set full captures of the main kernels (one launch each) from scripts/prof_step.py
set -u
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 python scripts/prof_step.py > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
}
cap up_halo32_h16 conv_up_halo_ts_kernel 2
cap wgrad32_h16 conv_wgrad32_tc_kernel 0
cap down32_h16 conv_down32_ts_kernel 0
cap up_c2i conv_up_c2i_kernel 0
cap linear_nt linear_nt_tc_kernel 0
cap btcvae_fwd3 btcvae_fwd3_kernel 0
ls -la gpurun_out/*.ncu-rep

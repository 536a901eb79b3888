#!/bin/bash
# what dp_init's MSM tuning (DP_MSM_TUNE=1) finds on this GPU: plain pipeline vs two batched-affine tree levels
export DP_MSM_TUNE=1
mkdir -p gpurun_out
for cfg in "1 20" "1 22" "8 22" "2 22"; do
    set -- $cfg
    python -m distributed_plonk_b200.tune 0 0 $1 $2 > gpurun_out/r02i_tune_w$1_2p$2.json 2> gpurun_out/r02i_tune_w$1_2p$2.err
    echo "W=$1 2^$2: $(cat gpurun_out/r02i_tune_w$1_2p$2.json) rc=$?"
done

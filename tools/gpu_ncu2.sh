#!/usr/bin/env bash
# ncu evidence, final code of the round: launch list of one forward + --set full for the top kernels (eval and training).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
MODE=${1:-tf32x3}
T=r01d
B="python bench.py --mode $MODE --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-alt"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${T}_$MODE.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dsconv_fused -s 40 -c 3 -o gpurun_out/prof_${T}_dsconv_$MODE -f $B > gpurun_out/ncu_ds.log 2>&1
echo "dsconv full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pw1x1_tc_kernel -s 40 -c 2 -o gpurun_out/prof_${T}_pw_$MODE -f $B > gpurun_out/ncu_pw.log 2>&1
echo "pw full rc=$?"
SMAAT_FUSE_DS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dw3x3_kernel -s 64 -c 2 -o gpurun_out/prof_${T}_dw_unfused -f $B > gpurun_out/ncu_dw.log 2>&1
echo "dw full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:cbam_|upsample2x" -s 110 -c 12 -o gpurun_out/prof_${T}_cbam_up -f $B > gpurun_out/ncu_cbam.log 2>&1
echo "cbam+upsample full rc=$?"
timeout 900 ncu --set full --clock-control none -k "regex:wgrad|dw3x3_bwd|bn_act_bwd|cbam_bwd" -s 60 -c 10 -o gpurun_out/prof_${T}_train -f python tools/prof_train.py > gpurun_out/ncu_train.log 2>&1
echo "train full rc=$?"
ls -la gpurun_out/prof_${T}_*.ncu-rep gpurun_out/launches_${T}_$MODE.csv

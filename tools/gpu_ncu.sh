#!/usr/bin/env bash
# ncu evidence for the round: launch list of one forward + --set full for the top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
MODE=${1:-tf32x3}
B="python bench.py --mode $MODE --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-alt"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01_$MODE.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# the eager warm-up forwards come first: skip them so the captured launch is warm-cache steady state
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dsconv_fused -s 40 -c 2 -o gpurun_out/prof_r01_dsconv_$MODE -f $B > gpurun_out/ncu_ds.log 2>&1
echo "dsconv full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pw1x1_tc_kernel -s 40 -c 2 -o gpurun_out/prof_r01_pw_$MODE -f $B > gpurun_out/ncu_pw.log 2>&1
echo "pw full rc=$?"
SMAAT_FUSE_DS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dw3x3_kernel -s 64 -c 2 -o gpurun_out/prof_r01_dw_unfused -f $B > gpurun_out/ncu_dw.log 2>&1
echo "dw full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cbam_ -s 100 -c 5 -o gpurun_out/prof_r01_cbam -f $B > gpurun_out/ncu_cbam.log 2>&1
echo "cbam full rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r01_$MODE.csv

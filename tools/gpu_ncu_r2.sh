#!/usr/bin/env bash
# (reports are kept small -- no --import-source -- because gpurun copies back at most 64 MiB)
# round 2 ncu evidence: launch list of one forward + --set full captures of the top kernels (1 GPU; numbers printed under ncu are never bench values)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
T=${1:-r02}
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-alt --no-train"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${T}.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# the fused DS kernels of one forward in steady state (skip the warm-up forwards' launches)
timeout 1200 ncu --set full --clock-control none -k regex:dsconv_tmem -s 36 -c 7 -o gpurun_out/prof_${T}_dsconv -f $B > gpurun_out/ncu_ds.log 2>&1
echo "dsconv_tmem full rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:pw1x1_tc_kernel -s 18 -c 3 -o gpurun_out/prof_${T}_pw -f $B > gpurun_out/ncu_pw.log 2>&1
echo "pw full rc=$?"
timeout 900 ncu --set full --clock-control none -k "regex:cbam_|upsample2x" -s 36 -c 9 -o gpurun_out/prof_${T}_cbam_up_dw -f $B > gpurun_out/ncu_cbam.log 2>&1
echo "cbam+upsample+dw full rc=$?"
SMAAT_FUSE_DS=0 timeout 900 ncu --set full --clock-control none -k regex:dw3x3_kernel -s 64 -c 1 -o gpurun_out/prof_${T}_dw_unfused -f $B > gpurun_out/ncu_dw.log 2>&1
echo "dw (unfused pass) full rc=$?"
ls -la gpurun_out/prof_${T}_*.ncu-rep gpurun_out/launches_${T}.csv

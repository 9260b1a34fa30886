#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
m=${1:-tf32x3}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dsconv_fused -s 2 -c 1 -o gpurun_out/prof_r01b_fused_$m -f python tools/prof_one.py dsconv $m 64 288 64 > gpurun_out/prof_r01b_fused_$m.log 2>&1
echo "$m rc=$?"

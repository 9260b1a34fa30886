#!/usr/bin/env bash
# launch list of one forward of the FINAL code (46 launches)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-alt --no-train"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$? $(wc -l < gpurun_out/launches_r02.csv) lines"

#!/usr/bin/env bash
# re-capture after the DS-kernel rewrite: launch list of one forward + --set full of the fused DS kernels (the other captures of
# tools/gpu_ncu_r2.sh are still current: those kernels did not change)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
T=${1:-r02}
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-alt --no-train"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${T}.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:dsconv_tmem -s 36 -c 7 -o gpurun_out/prof_${T}_dsconv -f $B > gpurun_out/ncu_ds.log 2>&1
echo "dsconv_tmem full rc=$?"
ls -la gpurun_out/prof_${T}_*.ncu-rep gpurun_out/launches_${T}.csv

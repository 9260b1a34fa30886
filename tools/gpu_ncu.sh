#!/usr/bin/env bash
# ncu evidence: (1) launch list of one bench step, (2) --set full for the dw and pw kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
MODE=${1:-tf32x3}
# launch list: warm-up = InferenceSession warm-up (2 fwd) + capture; profile the eager roofline pass instead (--no-graph)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$MODE.csv \
  python bench.py --mode $MODE --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_bench_$MODE.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dw3x3_kernel -s 16 -c 2 -o gpurun_out/prof_dw \
  python bench.py --mode $MODE --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_dw.log 2>&1
echo "dw full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pw1x1_tc_kernel -s 16 -c 3 -o gpurun_out/prof_pw_$MODE \
  python bench.py --mode $MODE --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_pw.log 2>&1
echo "pw full rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_$MODE.csv

#!/usr/bin/env bash
# event traces of CTA 0 (who waits for whom), single and dual issuer; per-layer times of the knock-outs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for cfg in "128 288 64" "256 144 128"; do
  for fl in 0 64; do SMAAT_DT_FLAGS=$fl timeout 120 python tools/dt_trace.py $cfg 64 18 2>&1 | tail -24; done
done
SMAAT_DT_FLAGS=16 timeout 120 python tools/dt_trace.py 128 288 64 64 18 2>&1 | tail -24
SMAAT_DT_FLAGS=4 timeout 120 python tools/dt_trace.py 128 288 64 64 18 2>&1 | tail -24
} > gpurun_out/dt_trace_r02.txt 2>&1
cat gpurun_out/dt_trace_r02.txt
for fl in 0 64 2 4 8 16; do
  echo "== SMAAT_DT_FLAGS=$fl"
  SMAAT_DT_FLAGS=$fl timeout 120 python tools/time_ds.py tf32x3 tmem 2>&1 | awk '{printf "%s ", $(NF-5)} END {print ""}'
done

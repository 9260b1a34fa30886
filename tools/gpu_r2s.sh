#!/usr/bin/env bash
# event traces of the lean issue loop (instrumented build): dual issuers (default) and single (flag 64)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for cfg in "128 288 64" "256 144 128"; do
  for fl in 0 64; do SMAAT_DT_FLAGS=$fl timeout 120 python tools/dt_trace.py $cfg 64 18 2>&1 | tail -24; done
done
} > gpurun_out/dt_trace_r02b.txt 2>&1
cut -c1-200 gpurun_out/dt_trace_r02b.txt

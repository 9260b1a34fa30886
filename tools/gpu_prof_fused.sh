#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
m=${1:-tf32}
timeout 600 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:dsconv_fused -s 2 -c 1 -o gpurun_out/prof_fused2_$m -f python tools/prof_one.py dsconv $m > gpurun_out/prof_fused2_$m.log 2>&1
echo "$m rc=$?"

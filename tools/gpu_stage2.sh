#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 tools/probe.bin dw > gpurun_out/probe.log 2>&1
timeout 120 tools/probe.bin pw 1 | grep -v "y\[" >> gpurun_out/probe.log 2>&1
cat gpurun_out/probe.log
rm -f gpurun_out/summary.log
bash tools/gpu_stage1.sh

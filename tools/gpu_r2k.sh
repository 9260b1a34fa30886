#!/usr/bin/env bash
cd "$(dirname "$0")/.."
bash tools/gpu_ncu_r2.sh r02 2>&1 | tail -10
du -sh gpurun_out

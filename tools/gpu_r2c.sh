#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_api_paths.py -q -m gpu -p no:cacheprovider -s -k "gradients_at_288" 2>&1 | grep -E "passed|failed|dL/dx|worst|Error" | cut -c1-250
for a in "12 288 64" "64 288 64" "128 288 64" "64 144 128" "256 144 128"; do timeout 120 python tools/dt_timing.py $a tf32x3 2>&1 | tail -6; done
timeout 120 python tools/dt_timing.py 128 288 64 tf32 2>&1 | tail -6

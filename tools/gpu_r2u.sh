#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "dsconv" > gpurun_out/pytest_r2u.log 2>&1
echo "dsconv parity rc=$? $(tail -n 1 gpurun_out/pytest_r2u.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r2u.log | cut -c1-200 | head -20

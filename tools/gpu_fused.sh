#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider -rf --tb=short tests/test_gpu_kernels.py -k "dsconv" > gpurun_out/fused_tests.log 2>&1
echo "fused tests rc=$? $(tail -n 1 gpurun_out/fused_tests.log)"
grep -E "^(FAILED|E  )" gpurun_out/fused_tests.log | head -30

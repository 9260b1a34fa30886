#!/usr/bin/env bash
# full -m gpu suite exactly as the driver runs it + smoke + default bench + training bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|E  )" gpurun_out/pytest_gpu.log | cut -c1-200 | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
MODES="tf32 tf32x3" bash tools/gpu_bench.sh 2>&1 | grep -E "fps|smaat_" | grep -v "^#" 
timeout 600 python bench_train.py --steps 5 --warmup 2 2>&1 | tail -1

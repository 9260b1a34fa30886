#!/usr/bin/env bash
# extras on the final code: configs[4] (576x576, B=8) and the reference arm of bench.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/bench_576.py > gpurun_out/bench_576_r02.jsonl 2>&1; cat gpurun_out/bench_576_r02.jsonl | cut -c1-300
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_r02.log 2> gpurun_out/bench_reference_r02.err; echo "reference arm rc=$?"; cut -c1-900 gpurun_out/bench_reference_r02.log

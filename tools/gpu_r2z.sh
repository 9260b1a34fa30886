#!/usr/bin/env bash
# the intermittent launch failure on 576x576 inputs in tf32 mode: repeat the forward many times after the in_full barrier fix
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "dsconv" > gpurun_out/pytest_r2z.log 2>&1; echo "dsconv parity rc=$? $(tail -n 1 gpurun_out/pytest_r2z.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_r2z.log | cut -c1-200 | head -5
for rep in 1 2 3 4; do for fl in 0 64; do
  echo -n "rep $rep flags $fl: "; SMAAT_DT_FLAGS=$fl timeout 100 python tools/dbg_576.py tf32 8 2>&1 | grep -v "Warning\|^$" | head -4 | tr '\n' ' ' | cut -c1-300; echo
done; done
timeout 200 python tools/bench_576.py 2>&1 | grep -v Warning | head -4 | cut -c1-300

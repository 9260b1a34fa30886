#!/usr/bin/env bash
# First on-GPU pass: every stage under its own timeout, logs into gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.log
  timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc  $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.log
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.log 2>&1
PT="python -m pytest -q -m gpu -rf --tb=short -p no:cacheprovider"
run k_nonpw   400 $PT tests/test_gpu_kernels.py -k "not pw1x1"
run k_pw_fp32 400 $PT tests/test_gpu_kernels.py -k "pw1x1 and fp32"
run k_pw_tf32 400 $PT tests/test_gpu_kernels.py -k "pw1x1 and tf32 and not tf32x3"
run k_pw_x3   400 $PT tests/test_gpu_kernels.py -k "pw1x1 and tf32x3"
run m_fp32    400 $PT tests/test_gpu_modules.py -k "fp32 or standalone or train_mode"
run m_tf32    400 $PT tests/test_gpu_modules.py -k "tf32"
run smoke     300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_fp32 600 python bench.py --mode fp32 --steps 5 --warmup 3 --no-cpu-baseline
run bench_tf32 600 python bench.py --mode tf32 --steps 10 --warmup 3 --no-cpu-baseline
run bench_x3   600 python bench.py --mode tf32x3 --steps 10 --warmup 3
run full      900 $PT tests/test_gpu_full.py
cat gpurun_out/summary.log

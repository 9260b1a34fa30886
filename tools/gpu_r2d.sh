#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "dsconv" > gpurun_out/pytest_r2d_ds.log 2>&1; rc=$?
echo "dsconv kernel tests rc=$rc $(tail -n 1 gpurun_out/pytest_r2d_ds.log)"
grep -E "^(FAILED|E  )" gpurun_out/pytest_r2d_ds.log | cut -c1-240 | head -20
[ $rc -ne 0 ] && exit 0
for a in "12 288 64" "64 288 64" "128 288 64" "64 144 128" "256 144 128"; do timeout 120 python tools/dt_timing.py $a tf32x3 2>&1 | tail -6; done
SMAAT_BENCH_LAYERS=1 timeout 600 python bench.py --no-cpu-baseline --no-alt > gpurun_out/bench_r2d.log 2>gpurun_out/bench_r2d.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2d.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks')})
print('roofline', {k:d['roofline'][k] for k in ('kernel','bound','frac','frac_hbm','frac_tensor','ms_per_step')})
PY
grep "^# smaat_dsconv" gpurun_out/bench_r2d.err | sort -u

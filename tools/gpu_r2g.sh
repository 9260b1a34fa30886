#!/usr/bin/env bash
# stager fix: parity of the fused kernel, stage timers, per-layer A/B, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "dsconv" -x > gpurun_out/pytest_r2g.log 2>&1; rc=$?
tail -2 gpurun_out/pytest_r2g.log
if [ $rc -ne 0 ]; then echo "parity failed or hung (rc=$rc): stopping"; grep -E "^(FAILED|E  )" gpurun_out/pytest_r2g.log | head -5; exit 1; fi
for a in "12 288 64" "64 288 64" "128 288 64" "64 144 128" "256 144 128" "256 72 256"; do timeout 120 python tools/dt_timing.py $a tf32x3 2>&1 | tail -5; done
echo "== tmem kernel"; timeout 120 python tools/time_ds.py tf32x3 tmem 2>&1 | tail -14
echo "== tmem kernel, tf32"; timeout 120 python tools/time_ds.py tf32 tmem 2>&1 | tail -14
SMAAT_BENCH_LAYERS=1 timeout 900 python bench.py --no-cpu-baseline --no-train > gpurun_out/bench_r2g.log 2>gpurun_out/bench_r2g.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2g.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print('e2e', d['e2e']['value'], 'via_api', d['via_reference_api']['value'], 'alt', d['alt_mode'])
print('roofline', {k:d['roofline'][k] for k in ('kernel','bound','frac','frac_hbm','frac_tensor','ms_per_step')})
for k,v in d['kernels'].items(): print(f"   {k:28s} n={v['launches_per_step']:3d} {v['ms_per_step']:7.3f} ms ({100*v['frac_hbm']:5.1f}% hbm) {v['tflops']:6.1f} TF")
PY

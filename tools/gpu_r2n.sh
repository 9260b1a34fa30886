#!/usr/bin/env bash
# diagnostics: per-CTA finish spread of the TMEM-operand DS kernel, per-kernel breakdown of the training step, recompute option
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -x -q -m gpu -p no:cacheprovider -k "upsample or recompute or dsconv or abi" > gpurun_out/pytest_r2n.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_r2n.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_r2n.log | cut -c1-220 | head
for cfg in "128 288 64" "64 288 64" "256 144 128" "128 144 128" "512 72 256" "12 288 64"; do
  timeout 120 python tools/dt_timing.py $cfg 2>&1 | grep -v "^  tma\|^  producer\|^  epilogue"
done
timeout 300 python tools/train_breakdown.py 32 2>&1 | tail -50
SMAAT_BENCH_LAYERS=1 timeout 420 python bench.py > gpurun_out/bench_r2n.log 2> gpurun_out/bench_r2n.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r2n.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
    print('train', json.dumps(d['train'])[:900])
    for k,v in d['kernels'].items(): print(f"   {k:28s} n={v['launches_per_step']:3d} {v['ms_per_step']:7.3f} ms ({100*v['frac_hbm']:5.1f}% hbm) {v['tflops']:6.1f} TF")
except Exception as e: print('no bench line:', e)
PY
grep "^# smaat" gpurun_out/bench_r2n.err | grep "cbam\|upsample" | sort -u | head -40

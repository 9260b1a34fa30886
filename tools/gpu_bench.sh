#!/usr/bin/env bash
# bench only (per-layer table on stderr); MODES="tf32 tf32x3"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in ${MODES:-tf32 tf32x3}; do
  SMAAT_BENCH_LAYERS=1 timeout 600 python bench.py --mode $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$m.log 2>&1
  echo "== $m rc=$?"; grep -h '^# ' gpurun_out/bench_$m.log
  python - "$m" <<'PY'
import json,sys
m=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/bench_{m}.log').read().strip().splitlines()[-1])
    print(f"{m}: value={d['value']:.0f} fps  ms/step={d['ms_per_step']:.2f}  e2e={d['e2e']['value']:.0f}  clocks={d['clocks']}")
    for k,v in d['kernels'].items(): print(f"   {k:26s} n={v['launches_per_step']:3d} {v['ms_per_step']:7.3f} ms  {v['achieved_GBps']:7.0f} GB/s ({100*v['frac_hbm']:5.1f}% hbm) {v['tflops']:6.1f} TF")
except Exception as e:
    print('no json', e); print(open(f'gpurun_out/bench_{m}.log').read()[-1500:])
PY
done

#!/usr/bin/env bash
# Quick regression + bench; prints a compact summary only.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/summary.log
run() { local name=$1 to=$2; shift 2; timeout -k 10 "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "$name rc=$? $(tail -n 1 gpurun_out/$name.log | cut -c1-150)" >> gpurun_out/summary.log; }
PT="python -m pytest -q -m gpu -rf --tb=short -p no:cacheprovider"
run kernels 600 $PT tests/test_gpu_kernels.py
run modules 600 $PT tests/test_gpu_modules.py
run full 900 $PT tests/test_gpu_full.py
for m in ${MODES:-tf32 tf32x3}; do
  SMAAT_BENCH_LAYERS=1 run bench_$m 600 python bench.py --mode $m --steps 10 --warmup 3 --no-cpu-baseline
done
cat gpurun_out/summary.log
grep -h '^# ' gpurun_out/bench_*.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, 'no json'); continue
    print(f, f"value={d['value']:.0f} fps  ms/step={d['ms_per_step']:.2f}  e2e={d['e2e']['value']:.0f}  clocks={d['clocks']}")
    for k,v in d['kernels'].items(): print(f"   {k:26s} n={v['launches_per_step']:3d} {v['ms_per_step']:7.3f} ms  {v['achieved_GBps']:7.0f} GB/s ({100*v['frac_hbm']:5.1f}% hbm) {v['tflops']:6.1f} TF")
PY

#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python bench.py --no-train --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_r3c.log 2> gpurun_out/bench_r3c.err; echo "bench rc=$? lines=$(wc -l < gpurun_out/bench_r3c.log)"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3c.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('cbam', d['cbam_roofline']); print('roof', d['roofline']['frac'], d['depthwise_roofline']['frac'])
PY
tail -3 gpurun_out/bench_r3c.err | cut -c1-200

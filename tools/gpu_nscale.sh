#!/usr/bin/env bash
# multi-GPU bench exactly as the driver launches it (torchrun, one rank per GPU): usage gpu_nscale.sh N
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
grep -E "^\[nccl\]" gpurun_out/bench_n$N.err | head -8 | cut -c1-220
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','clocks')})
print('e2e', d['e2e']['value'], 'via_api', d['via_reference_api']['value'])
print('train', json.dumps(d['train'], indent=1)[:3000])
PY
tail -3 gpurun_out/bench_n$N.err | cut -c1-300

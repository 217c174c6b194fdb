#!/bin/bash
# quick A/B on the GPU: headline tick time (kernel split) at 65,536 envs and, optionally, other batch sizes
# usage: scripts/ab.sh TAG [envs ...]
TAG=${1:-ab}; shift || true
mkdir -p gpurun_out
for n in 65536 "$@"; do
  python bench.py --envs $n --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --rollout-k 0 > gpurun_out/${TAG}_${n}.json 2> gpurun_out/${TAG}_${n}.err
  python - "$TAG" "$n" <<'PY'
import json,sys
tag,n=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f'gpurun_out/{tag}_{n}.json').read().strip().splitlines()[-1])
    print(tag, n, 'value=%.4g'%d['value'], 'ms=%.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['kernel_ms_per_tick'].items()})
except Exception as e:
    print(tag, n, 'ERR', e, open(f'gpurun_out/{tag}_{n}.err').read()[-400:])
PY
done

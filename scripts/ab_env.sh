#!/bin/bash
# A/B of execution-strategy switches on the GPU: scripts/ab_env.sh TAG "VAR=val VAR=val" ["VAR=val ..." ...]  (65,536 envs)
TAG=$1; shift
mkdir -p gpurun_out
for v in "" "$@"; do
  env $v python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --rollout-k 0 > gpurun_out/${TAG}_tmp.json 2> gpurun_out/${TAG}_tmp.err
  python - "$v" gpurun_out/${TAG}_tmp.json <<'PY' | tee -a gpurun_out/${TAG}.txt
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print('%-60s value=%.4g ms=%.4f' % (sys.argv[1] or '(default)', d['value'], d['ms_per_step']), {k:round(v,4) for k,v in d['kernel_ms_per_tick'].items() if v > 0.01})
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done

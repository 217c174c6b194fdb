#!/bin/bash
# quick GPU pass: GPU tests + headline bench (+ optional extra configs given as arguments)
set -u
TAG=${1:-r2x}; shift || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log
for cfg in normal65k "$@"; do
  timeout 600 python bench.py --config $cfg --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_${cfg}.json 2> gpurun_out/${TAG}_bench_${cfg}.err; echo "bench $cfg rc=$?"
done
python - <<'PY'
import json,glob,sys,os
tag=os.environ.get("TAG","")
for f in sorted(glob.glob('gpurun_out/%s_bench_*.json' % sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/*_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value=%.4g'%d['value'], 'ms=%.4f'%d['ms_per_step'], 'e2e=%.4g'%((d.get('e2e') or {}).get('value',0)), 'rollout=%.4g'%((d.get('rollout') or {}).get('value',0)), {k:round(v,4) for k,v in (d.get('kernel_ms_per_tick') or {}).items()}, (d.get('e2e') or {}).get('replay_matches_recording'))
    except Exception as e:
        print(f, 'ERR', e)
PY

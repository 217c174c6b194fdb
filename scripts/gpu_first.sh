#!/bin/bash
# first GPU pass of a build: external probes, smoke, GPU tests, one bench line per BASELINE config
set -u
mkdir -p gpurun_out
TAG=${1:-r2a}
python -c "import Box2D, gymnasium; print('Box2D', Box2D.__version__, 'gymnasium', gymnasium.__version__)" > gpurun_out/${TAG}_box2d_probe.txt 2>&1
echo "rc=$?" >> gpurun_out/${TAG}_box2d_probe.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/${TAG}_smi.txt
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${TAG}_pytest.log
for cfg in normal65k shooting4k defense65k defense65k_weak normal1M actor262k; do
  timeout 600 python bench.py --config $cfg --steps 200 --warmup 20 > gpurun_out/${TAG}_bench_${cfg}.json 2> gpurun_out/${TAG}_bench_${cfg}.err; echo "bench $cfg rc=$?"
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_driver_args.json 2>&1; echo "bench driver-args rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2>&1; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/*_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'value=%.4g'%d['value'], 'ms=%.4f'%d['ms_per_step'], 'e2e=%.4g'%((d.get('e2e') or {}).get('value',0)), 'rollout=%.4g'%((d.get('rollout') or {}).get('value',0)), d.get('kernel_ms_per_tick'), d.get('episode_stats',{}).get('toi_events_per_step'))
    except Exception as e:
        print(f, 'ERR', e)
PY

python __graft_entry__.py --smoke 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --impl reference --steps 100 --warmup 50 2>/dev/null | tail -1 | cut -c1-300
python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_default.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_default.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['steps'], d['warmup'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'])"

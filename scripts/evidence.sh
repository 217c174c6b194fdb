#!/bin/bash
# One gpurun call that produces the evidence set of a build: usage scripts/evidence.sh <tag>  (files land in gpurun_out/)
T=${1:-r1c}; O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/pytest_$T.txt; cat $O/pytest_$T.txt
python scripts/parity_sweep.py 8192 400 > $O/parity_sweep_$T.txt 2>&1; tail -1 $O/parity_sweep_$T.txt
python bench.py --impl reference --steps 400 --warmup 300 > $O/bench_${T}_ref.json 2> $O/bench_${T}_ref.err
python bench.py --steps 400 --warmup 300 > $O/bench_${T}_final.json 2> $O/bench_${T}_final.err
python bench.py --envs 32768 --steps 200 --warmup 300 --no-cpu-baseline > $O/bench_${T}_32k.json 2>/dev/null
python bench.py --envs 131072 --steps 200 --warmup 300 --no-cpu-baseline > $O/bench_${T}_131k.json 2>/dev/null
python bench.py --envs 1048576 --steps 100 --warmup 300 --no-cpu-baseline > $O/bench_${T}_1M.json 2>/dev/null
for f in final 32k 131k 1M ref; do python - $O/bench_${T}_$f.json $f <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], d.get("value"), d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("gpu_launches"), d.get("cpu_baseline", {}).get("value"))
PY
done
python scripts/phase_cycles.py 65536 > $O/phase_cycles_$T.txt 2>&1; cat $O/phase_cycles_$T.txt
CMD="python bench.py --steps 20 --warmup 100 --no-cpu-baseline --e2e-steps 5"
$CMD > $O/plain_$T.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file $O/${T}_launches.csv $CMD > $O/ncu_a_$T.log 2>&1
$CMD > $O/plain_$T.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_general|k_fast' -s 220 -c 2 -f -o $O/${T}_full $CMD > $O/ncu_b_$T.log 2>&1
CMD2="python bench.py --envs 1048576 --steps 6 --warmup 60 --no-cpu-baseline --e2e-steps 3"
$CMD2 > $O/plain_${T}_1M.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 18 --csv --log-file $O/${T}_launches_1M.csv $CMD2 > $O/ncu_c_$T.log 2>&1
tail -2 $O/ncu_b_$T.log

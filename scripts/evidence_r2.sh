#!/bin/bash
# One gpurun call that produces the round-2 evidence set of a build: usage scripts/evidence_r2.sh <tag>
# (files land in gpurun_out/; scripts/extract_profiles_r2.py turns them into the tracked files under profiles/)
T=${1:-r2}; O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${T}_pytest.txt; cat $O/${T}_pytest.txt
python bench.py --impl reference --steps 400 --warmup 20 > $O/${T}_bench_reference_arm.json 2> $O/${T}_ref.err
for cfg in normal65k shooting4k defense65k defense65k_weak normal1M actor262k; do
  python bench.py --config $cfg --steps 300 --warmup 20 > $O/${T}_bench_${cfg}.json 2> $O/${T}_bench_${cfg}.err; echo "bench $cfg rc=$?"
done
python bench.py --envs 131072 --steps 200 --warmup 20 --no-cpu-baseline > $O/${T}_bench_131k_envs.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/${T}_bench_driver_args.json 2>/dev/null
python scripts/phase_cycles.py 65536 > $O/${T}_phase_cycles.txt 2>&1
# ncu: launch list of the timed region, then one full capture of the two kernels of a steady-state tick (400-tick pre-roll = 800 launches)
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --rollout-k 0"
$CMD > $O/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_general|k_fast|k_touch' -s 800 -c 50 --csv --log-file $O/${T}_launches.csv $CMD > $O/${T}_ncu_a.log 2>&1
$CMD > $O/${T}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_general|k_fast' -s 820 -c 2 -f -o $O/${T}_full $CMD > $O/${T}_ncu_b.log 2>&1
tail -2 $O/${T}_ncu_b.log
python - $O $T <<'PY'
import json, sys, glob
o, t = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob(f"{o}/{t}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value=%.4g" % d["value"], "ms=%.4f" % d["ms_per_step"], "e2e=%.4g" % ((d.get("e2e") or {}).get("value", 0)),
              "cpu=%.4g" % ((d.get("cpu_baseline") or {}).get("value", 0)))
    except Exception as e:
        print(f, "ERR", e)
PY

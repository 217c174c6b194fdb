#!/bin/bash
# usage: scripts/ab_sweep.sh <envs> "<VAR=val VAR=val>" ["<VAR=val ...>" ...]  -> one line per environment setting
N=$1; shift
for CFG in "$@"; do
  env $CFG python bench.py --envs $N --steps ${STEPS:-100} --warmup ${WARMUP:-250} --no-cpu-baseline --e2e-steps 5 2>/dev/null | tail -1 > /tmp/_b.json
  python - "$N" "$CFG" <<'PY'
import sys, json
d = json.load(open('/tmp/_b.json'))
print("envs", sys.argv[1], "[%s]" % sys.argv[2], "steps/s %.4g  ms/tick %.3f  e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), flush=True)
PY
done

#!/bin/bash
# usage: scripts/sweep_bench.sh "<tiers list>" "<envs list>"  -> prints one line per combination
for T in $1; do for N in $2; do
  HK_TIERS=$T python bench.py --envs $N --steps ${STEPS:-100} --warmup ${WARMUP:-250} --no-cpu-baseline --e2e-steps 5 2>/dev/null | tail -1 > /tmp/_b.json
  python - "$T" "$N" <<'PY'
import sys, json
d = json.load(open('/tmp/_b.json'))
print("tiers", sys.argv[1], "envs", sys.argv[2], "steps/s %.4g  ms/tick %.3f  e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), flush=True)
PY
done; done

#!/bin/bash
# usage: scripts/sweep_lanes.sh <envs> "<tiers>" "<lanes1 list>" "<lanes2 list>"
for T in $2; do for L1 in $3; do for L2 in $4; do
  HK_TIERS=$T HK_LANES1=$L1 HK_LANES2=$L2 python bench.py --envs $1 --steps ${STEPS:-100} --warmup ${WARMUP:-250} --no-cpu-baseline --e2e-steps 5 2>/dev/null | tail -1 > /tmp/_b.json
  python - "$T" "$L1" "$L2" "$1" <<'PY'
import sys, json
d = json.load(open('/tmp/_b.json'))
print("envs", sys.argv[4], "tiers", sys.argv[1], "lanes1 2^%s lanes2 2^%s" % (sys.argv[2], sys.argv[3]), "steps/s %.4g  ms/tick %.3f" % (d["value"], d["ms_per_step"]), flush=True)
PY
done; done; done

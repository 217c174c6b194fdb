"""One-line summary of bench.py JSON lines: scripts/print_bench.py FILE..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n=%s" % d["config"].get("envs_per_gpu"), "value=%.4g" % d["value"], "ms=%.4f" % d["ms_per_step"],
              "e2e=%.4g" % ((d.get("e2e") or {}).get("value", 0)), "rollout=%.4g" % ((d.get("rollout") or {}).get("value", 0)),
              {k: round(v, 4) for k, v in (d.get("kernel_ms_per_tick") or {}).items() if v > 0.01})
    except Exception as e:
        print(f, "ERR", e)

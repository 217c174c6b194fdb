"""CPU-only parity sweep: the product's device code compiled for the host (tests/hostsim, kernel cascade emulated tier
by tier) against the oracle, larger than the pytest tier.  usage: python scripts/host_parity_sweep.py [envs] [ticks]"""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import hostsim_lib as H
import oracle_lib as O
from parity_util import state_mismatches, outputs_equal

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 400
names = {O.POL_WEAK: "weak", O.POL_STRONG: "strong", O.POL_RANDOM: "random", O.POL_ZERO: "zero"}
t0, total, bad_any = time.time(), 0, False
for mode, p1, p2 in ((0, 2, 2), (0, 1, 2), (0, 3, 3), (1, 3, 2), (1, 2, 3), (2, 2, 4), (2, 2, 1), (2, 3, 3)):
    o = O.OracleBatch(n, mode=mode, seed=300 + mode, env_id_offset=5 * 10 ** 9, n_threads=os.cpu_count() or 1)
    h = H.HostSimBatch(n, mode=mode, seed=300 + mode, env_id_offset=5 * 10 ** 9, fast=True)
    ok = True
    for t in range(ticks):
        ro = o.step(None, p1, p2, O.STEP_AUTORESET)
        rh = h.step(None, p1, p2, O.STEP_AUTORESET)
        if outputs_equal(ro, rh) != []:
            print("OUTPUT MISMATCH", mode, p1, p2, "tick", t, outputs_equal(ro, rh))
            ok = False
            break
        if t % 25 == 24:
            bad = state_mismatches(o.get_state(), h.get_state())
            if len(bad):
                print("STATE MISMATCH", mode, p1, p2, "tick", t, bad[:5].tolist())
                ok = False
                break
    bad_any |= not ok
    total += n * (t + 1)
    print(f"mode {mode} p1 {names[p1]:6s} p2 {names[p2]:6s} {'ok ' if ok else 'BAD'} tiers(fast/mid/long) {h.fast_counts()} touch {h.touch_count()}", flush=True)
print(f"{total / 1e6:.2f} M env-steps compared in {time.time() - t0:.0f} s")
sys.exit(1 if bad_any else 0)

"""Large GPU-vs-oracle parity sweep (beyond what tests/ runs by default): full canonical state compared every 50 ticks,
outputs every tick.  usage: python scripts/parity_sweep.py [envs] [ticks]"""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import hockey_env_b200 as hk
import oracle_lib as O
from parity_util import state_mismatches

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 600
names = {O.POL_WEAK: "weak", O.POL_STRONG: "strong", O.POL_RANDOM: "random", O.POL_ZERO: "zero"}
total, t0 = 0, time.time()
for tiers in ("2", "3"):
    os.environ["HK_TIERS"] = tiers
    for mode, p1, p2 in ((0, 2, 2), (0, 1, 2), (0, 3, 3), (1, 3, 2), (1, 2, 3), (2, 2, 4), (2, 2, 1), (2, 3, 3)):
        env = hk.HockeyVecEnv(n, mode=hk.Mode(mode), device="cuda:0", seed=int(os.environ.get("HK_SWEEP_SEED", "100")) + mode, env_id_offset=3 * 10 ** 9,
                              p1=names[p1], p2=names[p2], want_agent_two=True)
        ora = O.OracleBatch(n, mode=mode, seed=int(os.environ.get("HK_SWEEP_SEED", "100")) + mode, env_id_offset=3 * 10 ** 9, n_threads=os.cpu_count() or 1)
        ok = True
        for t in range(ticks):
            env.step()
            ro = ora.step(None, p1, p2, O.STEP_AUTORESET)
            if not (np.array_equal(env.obs.cpu().numpy(), ro["obs"]) and np.array_equal(env.done.cpu().numpy(), ro["done"])
                    and np.array_equal(env.reward.cpu().numpy(), ro["reward"].astype(np.float32))
                    and np.array_equal(env.info2.cpu().numpy(), ro["info2"].astype(np.float32))):
                print("OUTPUT MISMATCH", tiers, mode, p1, p2, "tick", t)
                ok = False
                break
            if t % 50 == 49:
                bad = state_mismatches(ora.get_state(), env.get_full_state().cpu().numpy().view(np.uint32))
                if len(bad):
                    print("STATE MISMATCH", tiers, mode, p1, p2, "tick", t, bad[:5].tolist())
                    ok = False
                    break
        s = env.stats()
        total += n * (t + 1)
        print(f"tiers {tiers} mode {mode} p1 {names[p1]:6s} p2 {names[p2]:6s} {'ok ' if ok else 'BAD'} episodes {int(s['episodes'])} "
              f"toi_events {int(s['toi_events'])} overflows {int(s['overflows'])}", flush=True)
print(f"{total / 1e6:.1f} M env-steps compared in {time.time() - t0:.0f} s")

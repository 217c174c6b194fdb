"""Prints the share of block-cycles the general tiers spend in each tick phase (diagnostics)."""
import ctypes as C
import os
import sys
if "--trace" in sys.argv:
    os.environ["HK_LANE_TRACE"] = "1"  # the per-lane TOI split is only collected when tracing
import torch
sys.path.insert(0, ".")
import hockey_env_b200 as hk

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = hk.HockeyVecEnv(n, device="cuda:0", seed=0, p1="strong", p2="strong")
env.reset(one_starting=(torch.arange(n, device="cuda:0") % 2).to(torch.int8))
for _ in range(300):
    env.step()
torch.cuda.synchronize()
out = (C.c_double * 8)()
hk._lib.check(env.L.hk_debug_phase_cycles(env._h, out))
v = list(out)
for t in range(2):
    tot = sum(v[4 * t:4 * t + 4]) or 1.0
    print(f"tier {t + 1}: " + "  ".join(f"{name} {100 * v[4 * t + k] / tot:5.1f}%" for k, name in enumerate(("policy+collide", "island-solve", "TOI", "finish"))), f" total {tot:.3g} block-cycles")
if "--trace" not in sys.argv:
    f = (C.c_double * 6)()
    hk._lib.check(env.L.hk_debug_finish_cycles(env._h, f))
    f = list(f)
    blocks = 300 * 134.0
    print("finish phase per block-tick (warp 0): " + "  ".join(f"{nm} {x / blocks:8.0f}" for nm, x in zip(
        ("pre", "commit", "tickFinish", "store", "flush", "final-barrier"), f)), " (cycles)")
tot1 = sum(v[0:4]) or 1.0
print(f"tier-1 TOI phase split (block max of per-lane cycles, tiers=2 only): evaluation calls {100 * v[4] / tot1:.1f}% of tier time, event handling {100 * v[5] / tot1:.1f}%")

"""GPU experiment: fused rollout kernel vs K x cascade (correctness on a small batch first, then throughput)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import hockey_env_b200 as hk
from parity_util import state_mismatches

def mk(n, fused, gw=None, seed=3, run=None, fill=None):
    os.environ["HK_FUSED"] = "1" if fused else "0"
    if gw: os.environ["HK_FUSED_GEN_WARPS"] = str(gw)
    if run: os.environ["HK_FUSED_FAST_RUN"] = str(run)
    if fill: os.environ["HK_FUSED_FILL"] = str(fill)
    e = hk.HockeyVecEnv(n, device="cuda:0", seed=seed, p1="strong", p2="strong")
    for k in ("HK_FUSED", "HK_FUSED_GEN_WARPS", "HK_FUSED_FAST_RUN", "HK_FUSED_FILL"): os.environ.pop(k, None)
    return e
st = lambda e: e.get_full_state().cpu().numpy().view(np.uint32)
for n, k in ((512, 8), (4096, 32), (65536, 64)):
    a, b = mk(n, True), mk(n, False)
    for r in range(3):
        a.rollout(k); b.rollout(k)
        torch.cuda.synchronize()
        bad = state_mismatches(st(a), st(b))
        print(f"n={n} k={k} call {r}: mismatches {len(bad)}", bad[:4].tolist() if len(bad) else "", flush=True)
    sa, sb = a.stats(), b.stats()
    print("  stats equal:", all(sa[x] == sb[x] for x in ("episodes", "wins", "env_steps", "toi_events", "velocity_iterations")), flush=True)
res = {}
import itertools
variants = [("cascade", False, None, None, None)] + [(f"fused_r{r}", True, None, r, None) for r in (1, 2, 3, 4, 8, 64)]
for n in (65536,):
    for name, fused, gw, run, fill in variants:
        if "excl" in name: os.environ["HK_FUSED_EXCLUSIVE"] = "1"
        if "nocarve" in name: os.environ["HK_CARVEOUT"] = "0"
        e = mk(n, fused, gw, run=run, fill=fill)
        os.environ.pop("HK_FUSED_EXCLUSIVE", None); os.environ.pop("HK_CARVEOUT", None)
        e.reset(one_starting=(torch.arange(n, device="cuda:0") % 2).to(torch.int8))
        e.rollout(400)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            e.rollout(64)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        import ctypes as C
        pc = (C.c_double * 8)()
        e.L.hk_debug_phase_cycles(e._h, pc)
        pc = list(pc)
        res[f"{n}_{name}"] = {"ms_per_tick": ms / 256, "env_steps_per_s": n * 256 / (ms * 1e-3)}
        if fused and pc[0] > 0:
            res[f"{n}_{name}"].update({"rounds_per_block_tick": pc[0] / 148 / 656, "fill": pc[1] / pc[0], "round_cycles": pc[2] / pc[0],
                                       "fast_batches_per_block_tick": pc[4] / 148 / 656, "fast_phases_per_block_tick": pc[5] / 148 / 656,
                                       "fast_phase_cycles": pc[6] / max(pc[5], 1)})
        print(n, name, res[f"{n}_{name}"], flush=True)
        e.close()
print(json.dumps(res))

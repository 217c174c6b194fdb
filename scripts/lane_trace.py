"""Diagnostics: which warps / envs form the critical path of the general tier (HK_LANE_TRACE=1)."""
import os
import sys
os.environ["HK_LANE_TRACE"] = "1"
import numpy as np
import torch
sys.path.insert(0, ".")
import hockey_env_b200 as hk

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = hk.HockeyVecEnv(n, device="cuda:0", seed=0, p1="strong", p2="strong")
env.reset(one_starting=(torch.arange(n, device="cuda:0") % 2).to(torch.int8))
for _ in range(300):
    env.step()
nw = n // 32 + 8
names = ("policy+collide", "solve", "toi-tasks", "toi-events")
for rep in range(4):
    env.step()
    torch.cuda.synchronize()
    buf = np.zeros(20 * nw + 2 * n, np.uint32)
    hk._lib.check(env.L.hk_debug_lane_trace(env._h, buf.ctypes.data, buf.size))
    w = buf[:4 * nw].reshape(nw, 4).astype(np.int64)
    rec = buf[4 * nw:4 * nw + 2 * n].reshape(n, 2)
    blk = buf[4 * nw + 2 * n:].reshape(nw, 16).astype(np.int64)
    used = rec[:, 0] != 0
    r0 = rec[used, 0]
    gw = rec[used, 1]
    sweeps, toi, kind, shape, ab = r0 & 0xFFF, (r0 >> 12) & 0xF, (r0 >> 16) & 0xF, (r0 >> 20) & 0xFF, r0 >> 31
    print(f"tick {rep}: general-tier envs {used.sum()}  warps with work {(w.sum(1) > 0).sum()}  aborted {ab.sum()}")
    tot = w.sum(1)
    print("  warp cycles: mean", int(tot[tot > 0].mean()), "p50", int(np.percentile(tot[tot > 0], 50)), "p99", int(np.percentile(tot[tot > 0], 99)), "max", int(tot.max()))
    for k in range(4):
        c = w[:, k]
        print(f"  phase {names[k]:15s} mean {int(c[tot > 0].mean()):8d}  p99 {int(np.percentile(c[tot > 0], 99)):8d}  max {int(c.max()):8d}")
    for k in (1, 3):
        top = np.argsort(-w[:, k])[:5]
        for t in top:
            m = gw == t
            print(f"    slow {names[k]} warp {t}: {w[t].tolist()} lanes {m.sum()} sweeps {sorted(sweeps[m].tolist())[-4:]} toi {sorted(toi[m].tolist())[-4:]} "
                  f"shapes(nvc|pts<<4) {sorted(set(shape[m].tolist()))} kinds {sorted(set(kind[m].tolist()))}")
    cn = ("puck-racket", "racket-static", "puck-static/sensor", "other")
    for c in range(4):
        m = kind == c
        if not m.any():
            continue
        ws = np.unique(gw[m])
        print(f"  class {c} {cn[c]:18s}: envs {m.sum():6d} warps {len(ws):4d}  warp cycles mean " + " ".join(f"{names[k]} {int(w[ws, k].mean())}" for k in range(4)) +
              f"  total mean {int(tot[ws].mean())} max {int(tot[ws].max())}  events/env {toi[m].mean():.2f} sweeps/env {sweeps[m].mean():.1f}")
    hist = {}
    for sh, sw in zip(shape.tolist(), sweeps.tolist()):
        h = hist.setdefault(sh, [0, 0, 0])
        h[0] += 1
        h[1] += sw
        h[2] += sw >= 180
    print("  solve shapes (contacts|points<<4: envs, mean sweeps, envs with >=180 sweeps): " +
          "  ".join(f"{k & 15}c{k >> 4}p: {v[0]}, {v[1] / v[0]:.1f}, {v[2]}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1][0])))
    # block-level critical path: thread 0's stamps at the phase barriers
    bn = ("collide", "isl-begin", "vel-pool", "isl-end", "toi-eval", "toi-events", "finish")
    work = blk[:, 9] > 0
    order = np.argsort(-blk[:, 9])
    busy = blk[work & (blk[:, 9] > 20000)]
    print(f"  blocks with work {len(busy)}: total cycles mean {int(busy[:, 9].mean())} max {int(busy[:, 9].max())}; end-time spread {int(busy[:, 8].max() - busy[:, 8].min())} ns")
    print("  mean share per sub-phase: " + "  ".join(f"{bn[k]} {100 * busy[:, k].sum() / busy[:, 9].sum():.1f}%" for k in range(7)))
    for b in order[:6]:
        print(f"    slow block {b} (sm {blk[b, 7]}): total {blk[b, 9]}  " + "  ".join(f"{bn[k]} {blk[b, k]}" for k in range(7)) + f"  slowest solve unit: type {blk[b, 11]} {blk[b, 10]} cycles")
    for lab, m in (("no env done", busy[:, 15] == 0), ("some env done", busy[:, 15] > 0)):
        if m.any():
            print(f"    finish phase, blocks with {lab} ({m.sum()}): phase {int(busy[m, 6].mean())}  max-warp cycles commit {int(busy[m, 12].mean())} tickFinish {int(busy[m, 13].mean())} store+stats {int(busy[m, 14].mean())}")
    ut = ("-", "2c(1,1)", "2c(1,2)", "2c(2,1)", "2c(2,2)", "3c(1,1,1)", "3c(1,1,2)", "3c(1,2,1)", "3c(2,1,1)", "1c2p chunk", "1c1p chunk", "in place")
    for k in range(1, 12):
        m = busy[:, 11] == k
        if m.any():
            print(f"    slowest unit is {ut[k]:10s} in {m.sum():3d} blocks: cycles mean {int(busy[m, 10].mean())} max {int(busy[m, 10].max())}")

"""GPU experiment: cost of the host-facing output paths of one tick (65,536 NORMAL envs, p2 strong, host actions)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hockey_env_b200 as hk

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
res = {}
for name, zc in (("copy", False), ("zero_copy", True)):
    env = hk.HockeyVecEnv(n, device=dev, seed=1, p2="strong")
    env.reset(one_starting=(torch.arange(n, device=dev) % 2).to(torch.int8))
    rec = env.host_buffers()
    h_act = torch.empty((n, 4)).uniform_(-1, 1).pin_memory()
    for _ in range(400):
        env.step_host(h_act, rec, sync=False, zero_copy=zc)
    torch.cuda.synchronize()
    env.kernel_timing(True)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        env.step_host(h_act, rec, sync=True, zero_copy=zc)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 200 * 1e3
    kt, ks = env.kernel_times()
    res[name] = {"ms_per_tick_events": e0.elapsed_time(e1) / 200, "ms_per_tick_wall": wall,
                 "kernels_ms": {k: v / ks for k, v in kt.items()}, "env_steps_per_s": n * 200 / (e0.elapsed_time(e1) * 1e-3)}
    env.close()
# device-only step with external host-supplied actions already on the device (reference point)
env = hk.HockeyVecEnv(n, device=dev, seed=1, p2="strong")
env.reset(one_starting=(torch.arange(n, device=dev) % 2).to(torch.int8))
d_act = torch.empty((n, 4), device=dev).uniform_(-1, 1)
for _ in range(400):
    env.step(d_act)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    env.step(d_act)
e1.record()
torch.cuda.synchronize()
res["device_only"] = {"ms_per_tick_events": e0.elapsed_time(e1) / 200}
# raw copies
raw_h = torch.empty(n * 93, dtype=torch.uint8).pin_memory()
raw_d = torch.empty(n * 93, dtype=torch.uint8, device=dev)
for _ in range(5):
    raw_h.copy_(raw_d, non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(50):
    raw_h.copy_(raw_d, non_blocking=True)
e1.record()
torch.cuda.synchronize()
res["d2h_copy_ms"] = e0.elapsed_time(e1) / 50
res["d2h_GBps"] = n * 93 / (res["d2h_copy_ms"] * 1e-3) / 1e9
print(json.dumps(res, indent=1))

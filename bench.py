#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched HockeyEnv hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the CPU path (oracle restatement) on the host cores

A "step" is one tick of every env of the batch (one hk_step call per GPU = the kernel cascade k_fast -> k_general):
physics, contact sensing,
rewards, observation write, in-kernel BasicOpponents and auto-reset.  Workload (config.workload): 65,536
NORMAL-mode envs per GPU, strong-vs-strong BasicOpponent (weak scaling: envs are independent, sharded by
contiguous global env ids, no per-step communication; episode statistics are all-reduced once at the end).

Timing: W untimed warm-up steps, then K steps each bracketed by CUDA events on the launching stream with an L2
flush (256 MiB memset, untimed) between steps, barrier + synchronize on both sides, MAX over ranks.
"""
import argparse
import json
import os
import sys
import threading
import time

_ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _ROOT)

ENVS_PER_GPU = 65536
# algorithmic bytes per env-step of the dominant kernel (DESIGN.md "Data layout"): 256 B state read + 256 B state
# written + obs 72 + reward 4 + done 1 + info 16 (in-kernel opponents: no action read)
ALGO_BYTES_PER_ENV_STEP = 256 + 256 + 72 + 4 + 1 + 16
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (k_general<1>) for one launch at 65,536 envs, from the
# `ncu --set full` capture summarised in profiles/r1e_raw_metrics.csv (44.3 MB read + 33.8 MB written)
NCU_TRAFFIC_BYTES_PER_LAUNCH_65536 = 78_094_592
# smsp__inst_executed.sum of one tick at 65,536 envs from the same capture (k_fast 22.4 M + k_general 51.6 M warp
# instructions) and the issue ceiling of SURVEY.md section 8(d): 148 SMs x 4 schedulers x 1 warp-inst/clk x 1.965 GHz
NCU_WARP_INST_PER_TICK_65536 = 22_403_347 + 51_615_094
ISSUE_PEAK_WARP_INST_PER_S = 148 * 4 * 1.965e9


def _peaks():
    p = os.path.join(_ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def run_reference(args, rank, world):
    """The reference arm for this tier: the CPU implementation of the path (the oracle restatement of
    hockey_env.py + Box2D step; the reference's own pybox2d build is not installable here) on all host threads."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(_ROOT, "tests"))
    import oracle_lib as O
    cores = os.cpu_count() or 1
    n = 1024 * cores  # bounded sample of the 65,536-env workload (large enough to amortise the per-tick thread start)
    b = O.OracleBatch(n, mode=O.MODE_NORMAL, seed=args.seed, n_threads=cores)
    import numpy as np
    b.reset(one_starting=(np.arange(n) % 2).astype(np.int8))
    for _ in range(args.warmup):
        b.rollout(1, O.POL_STRONG, O.POL_STRONG)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.rollout(1, O.POL_STRONG, O.POL_STRONG)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{n} NORMAL envs x {args.steps} ticks, strong-vs-strong BasicOpponent, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "65536 NORMAL envs/GPU, strong-vs-strong BasicOpponent, auto-reset (bounded CPU sample)",
                   "sample_envs": n},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline(seed, budget_s=12.0):
    """Oracle ("port") timed on the host cores on a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(_ROOT, "tests"))
    import numpy as np
    import oracle_lib as O
    cores = os.cpu_count() or 1
    n = 128 * cores
    b = O.OracleBatch(n, mode=O.MODE_NORMAL, seed=seed, n_threads=cores)
    b.reset(one_starting=(np.arange(n) % 2).astype(np.int8))
    b.rollout(20, O.POL_STRONG, O.POL_STRONG)
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        b.rollout(50, O.POL_STRONG, O.POL_STRONG)
        steps += 50
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n} NORMAL envs x {steps} ticks, strong-vs-strong BasicOpponent, {cores} threads, {dt:.1f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=100)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import hockey_env_b200 as hk

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.envs
    env = hk.HockeyVecEnv(n, mode=hk.Mode.NORMAL, device=dev, seed=args.seed, env_id_offset=rank * n, p1="strong", p2="strong")
    env.reset(one_starting=(torch.arange(n, device=dev) % 2).to(torch.int8))
    flush = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        env.step()
    env.clear_stats()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        if flush is not None:
            flush.zero_()
        ev[k][0].record()
        env.step()
        ev[k][1].record()
    barrier()
    clocks = sampler.stop()
    kernel_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kernel_ms = float(t.item())
    value = world * n * args.steps / (kernel_ms * 1e-3)

    # ---- end to end through the public API with HOST buffers: player-1 actions come from pinned host memory every
    # tick, obs/reward/done/info go back to pinned host memory every tick (what a host-side agent would do).
    e2e_env = hk.HockeyVecEnv(n, mode=hk.Mode.NORMAL, device=dev, seed=args.seed + 1, env_id_offset=rank * n, p2="strong")
    h_act = torch.empty((n, 4), dtype=torch.float32).uniform_(-1, 1).pin_memory()
    d_act = torch.empty((n, 4), dtype=torch.float32, device=dev)
    h_obs = torch.empty((n, 18), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
    h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_info = torch.empty((n, 4), dtype=torch.float32).pin_memory()

    def e2e_tick():
        d_act.copy_(h_act, non_blocking=True)
        obs, rew, done, _, _ = e2e_env.step(d_act)
        h_obs.copy_(obs, non_blocking=True)
        h_rew.copy_(rew, non_blocking=True)
        h_done.copy_(done, non_blocking=True)
        h_info.copy_(e2e_env.info, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()  # the host needs the results before it can pick the next action

    for _ in range(max(3, args.warmup // 10)):
        e2e_tick()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        e2e_tick()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.e2e_steps / (float(t.item()) * 1e-3)

    # ---- end-of-run statistics: the only collective on this path (NCCL all-reduce of 16 doubles)
    st = env.stats_tensor()
    if world > 1:
        dist.all_reduce(st, op=dist.ReduceOp.SUM)
    st = st.cpu().tolist()

    if rank == 0:
        peak, peak_src = _peaks()
        bytes_per_launch = ALGO_BYTES_PER_ENV_STEP * n
        achieved = bytes_per_launch / (kernel_ms / args.steps * 1e-3) / 1e9
        out = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n} NORMAL envs per GPU, strong-vs-strong in-kernel BasicOpponent, auto-reset, all "
                                   "per-tick outputs (obs/reward/done/info) written", "envs_per_gpu": n,
                       "parallelism": f"env-sharded x{world}, no per-step communication",
                       "l2": "no flush" if flush is None else "flushed between timed steps (256 MiB memset, untimed)"},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": n * 16,
                    "d2h_bytes_per_step": n * (72 + 4 + 1 + 16), "steps": args.e2e_steps,
                    "note": "player-1 actions from pinned host memory, obs/reward/done/info to pinned host memory, "
                            "host sync every tick; player 2 = in-kernel strong BasicOpponent"},
            # kernels of this repo launched in the timed region: k_fast, k_touch and the general tier(s) per tick
            "gpu_launches": args.steps * env.launches_per_step(),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH_65536 if n == 65536 else None,
                         "traffic_unit": "bytes per launch of the dominant kernel k_general<1> (ncu, profiles/)",
                         "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP,
                         "issue": None if n != 65536 else {
                             "achieved": NCU_WARP_INST_PER_TICK_65536 / (kernel_ms / args.steps * 1e-3), "peak": ISSUE_PEAK_WARP_INST_PER_S,
                             "unit": "warp-inst/s", "frac": NCU_WARP_INST_PER_TICK_65536 / (kernel_ms / args.steps * 1e-3) / ISSUE_PEAK_WARP_INST_PER_S,
                             "note": "warp instructions per tick from ncu (profiles/r1e_raw_metrics.csv) / measured tick time"},
                         "note": "achieved = algorithmic bytes of one tick (all kernels of the cascade) / tick time; the path "
                                 "is instruction-fetch/divergence bound, not HBM-bound (DESIGN.md section 4, profiles/README.md)"},
            "episode_stats": {"episodes": st[0], "wins": st[1], "losses": st[2], "draws": st[3], "env_steps": st[4],
                              "mean_len": st[8] / max(st[0], 1), "velocity_iters_per_step": st[11] / max(st[4], 1),
                              "toi_events_per_step": st[12] / max(st[4], 1), "overflows": st[13]},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.seed)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

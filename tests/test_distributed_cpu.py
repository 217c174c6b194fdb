"""CPU tier, world_size 2 over gloo: the multi-GPU path's host logic -- contiguous global-env-id shards with no
per-tick communication, and the single end-of-run all-reduce(SUM) of the episode-statistics vector (bench.py).
Ranks step their shard with the host build of the device code (tests/hostsim); rank 0 checks the reduced
statistics against an unsharded run."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_per_rank, steps, q):
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hostsim_lib as H
    import oracle_lib as O
    env = H.HostSimBatch(n_per_rank, mode=0, seed=17, env_id_offset=rank * n_per_rank, fast=True)
    for _ in range(steps):
        env.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)      # no communication per tick
    st = torch.from_numpy(env.stats().copy())
    dist.barrier()
    dist.all_reduce(st, op=dist.ReduceOp.SUM)                               # the only collective of the path
    if rank == 0:
        q.put(st.numpy().tolist())
    dist.destroy_process_group()


def test_two_rank_shards_and_stats_allreduce(hostsim, oracle):
    world, n_per_rank, steps = 2, 48, 320
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per_rank, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    reduced = np.array(q.get(timeout=300))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = hostsim.HostSimBatch(world * n_per_rank, mode=0, seed=17, env_id_offset=0, fast=True)
    for _ in range(steps):
        ref.step(None, oracle.POL_STRONG, oracle.POL_STRONG, oracle.STEP_AUTORESET)
    want = ref.stats()
    assert reduced[4] == world * n_per_rank * steps
    for k in (0, 1, 2, 3, 4, 8, 9, 10, 12):            # episodes, W/L/D, steps, lengths, touches, TOI events
        assert reduced[k] == want[k], k
    assert reduced[5] == pytest.approx(want[5], rel=1e-9) and reduced[6] == pytest.approx(want[6], rel=1e-9)
    assert reduced[0] > 0

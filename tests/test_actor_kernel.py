"""The fused tensor-core actor kernel (csrc/hk_actor.cuh, hk_actor_forward) against the fp32 torch module it replaces
(ActorNetwork.forward, rl/td3/networks.py:17-20), with the reference's trained weights.  Tolerance: the kernel computes in
TF32 (layer 1) / bf16 (layers 2, 3) operands with fp32 accumulation and tanh.approx: mean |out - fp32| <= 4e-3; single
outputs can move by ~0.2 where the trained policy is steep (the same happens to the fp32 module under 1e-3 input noise),
so the check that matters is behavioural: the win rate against the strong BasicOpponent stays in the recorded interval."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "td3_actors.npz")


def test_param_block_layout_cpu():
    """The packed parameter block: element (n, k) of a weight sits where the UMMA descriptor expects it."""
    from hockey_env_b200.actor import FusedActor, load_td3_actor
    a = load_td3_actor(NPZ, device="cpu", name="stage_3")
    sd = {k: v.float() for k, v in a.state_dict().items()}
    blk = FusedActor.pack(sd)
    n1 = 256 * 24 * 4
    assert blk.dtype == torch.uint8 and blk.numel() == n1 + 256 * 256 * 2 + 16 * 256 * 2 + 1024 + 1024 + 64
    f1 = blk[:n1].view(torch.float32)                                           # layer 1: tf32 = f32 bits, 4 per 16-byte chunk
    w1 = sd["fc1.weight"]
    for n, k in ((0, 0), (5, 3), (255, 17), (77, 9)):
        assert f1[((k // 4) * (256 * 16) + n * 16 + (k % 4) * 4) // 4] == w1[n, k]
    assert f1[((18 // 4) * (256 * 16) + 3 * 16 + (18 % 4) * 4) // 4] == 0      # K padding
    bf2 = blk[n1:n1 + 256 * 256 * 2].view(torch.bfloat16)                       # layers 2, 3: bf16, 8 per chunk
    w2 = sd["fc2.weight"].to(torch.bfloat16)
    for n, k in ((0, 0), (200, 131), (255, 255)):
        assert bf2[((k // 8) * (256 * 16) + n * 16 + (k % 8) * 2) // 2] == w2[n, k]
    off3 = n1 + 256 * 256 * 2
    bf3 = blk[off3:off3 + 16 * 256 * 2].view(torch.bfloat16)
    w3 = sd["fc3.weight"].to(torch.bfloat16)
    assert bf3[((100 // 8) * (16 * 16) + 2 * 16 + (100 % 8) * 2) // 2] == w3[2, 100]
    assert bf3[((100 // 8) * (16 * 16) + 9 * 16 + (100 % 8) * 2) // 2] == 0       # N padding
    b1 = blk[off3 + 16 * 256 * 2:off3 + 16 * 256 * 2 + 1024].view(torch.float32)
    assert torch.equal(b1, sd["fc1.bias"])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 127, 128, 129, 4096, 100_003])
def test_fused_actor_matches_fp32_module(n):
    import hockey_env_b200 as hk
    ref = hk.load_td3_actor(NPZ, device="cuda:0", name="stage_3")
    fused = hk.FusedActor(ref)
    g = torch.Generator(device="cuda:0").manual_seed(n)
    # observations in the env's range: positions +-4, velocities +-10, angles, timers
    obs = torch.randn((n, 18), device="cuda:0", generator=g) * torch.tensor([2, 2, .5, 4, 4, 3, 2, 2, .5, 4, 4, 3, 3, 2, 8, 8, 4, 4], device="cuda:0")
    with torch.no_grad():
        want = ref(obs)
    got = fused(obs)
    torch.cuda.synchronize()
    assert got.shape == (n, 4) and torch.isfinite(got).all()
    err = (got - want).abs()
    # a trained policy has steep regions: the bound on single outputs is loose, the one on the mean is tight
    assert err.max().item() <= 0.35, err.max().item()
    assert err.mean().item() <= (4e-3 if n >= 4096 else 0.05), err.mean().item()   # the mean needs a sample to be one
    # writes into the first four columns of an [N, 8] action tensor, leaving the rest alone
    a8 = torch.full((n, 8), 7.0, device="cuda:0")
    fused(obs, out=a8)
    assert torch.equal(a8[:, :4], got) and (a8[:, 4:] == 7.0).all()


@pytest.mark.gpu
def test_fused_actor_plays_like_the_module():
    """Closed loop on real observations, and the recorded win rate (see test_actor_golden.py for the protocol)."""
    import hockey_env_b200 as hk
    from test_actor_golden import check_against_record, N_ENVS, QUOTA
    ref = hk.load_td3_actor(NPZ, device="cuda:0", name="stage_3")
    fused = hk.FusedActor(ref)
    env = hk.HockeyVecEnv(4096, device="cuda:0", seed=5, p2="strong")
    worst = 0.0
    obs = env.obs
    for _ in range(200):
        a = fused(obs)
        with torch.no_grad():
            worst = max(worst, (a - ref(obs)).abs().max().item())
        obs, *_ = env.step(a)
    assert worst <= 0.35, worst
    res = hk.evaluate(fused, n_episodes=N_ENVS * QUOTA, opponent="strong", num_envs=N_ENVS, seed=0)
    print("fused actor vs strong:", res)
    check_against_record(res, "stage_3", weak=False)

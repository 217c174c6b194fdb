"""Device-side TD3 update and prioritized replay against the reference's formulas
(rl/td3/learner.py:55-219, rl/replay/prioritized_buffer.py:6-69, rl/utils/torch_utils.py:12-24).  The modules are
device-agnostic torch code, so the arithmetic is checked on the CPU here; the GPU tier runs them behind the env."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _batch(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, 18, generator=g), torch.rand(n, 4, generator=g) * 2 - 1, torch.randn(n, generator=g),
            torch.randn(n, 18, generator=g), (torch.rand(n, generator=g) < 0.1).float())


def test_target_critic_loss_and_delayed_updates():
    from hockey_env_b200.td3 import DeviceTD3Learner, TD3Config, huber_weighted
    torch.manual_seed(0)
    cfg = TD3Config(target_action_noise_scale=0.0)           # noise off: the target is a closed formula
    L = DeviceTD3Learner(config=cfg, device="cpu", seed=1)
    s, a, r, s2, d = _batch(64)
    with torch.no_grad():
        q1, q2 = L.target_critic(s2, L.target_actor(s2).clamp(-1, 1))
        want = r + cfg.gamma * (1 - d) * torch.minimum(q1, q2)
    assert torch.allclose(L.compute_target(s2, r, d), want)
    # with noise: the perturbation of the target action is bounded by the clip
    L.cfg.target_action_noise_scale = 5.0
    t_noisy = L.compute_target(s2, r, d)
    assert torch.isfinite(t_noisy).all() and not torch.allclose(t_noisy, want)
    L.cfg.target_action_noise_scale = 0.2
    # smooth-L1: quadratic inside |d| < 1, linear outside, weighted, batch mean
    x, y = torch.tensor([0.0, 0.5, 3.0]), torch.tensor([0.0, 0.0, 0.0])
    w = torch.tensor([1.0, 2.0, 0.5])
    assert huber_weighted(x, y, w).item() == pytest.approx((0 + 0.5 * 2 * 0.25 + (3 - 0.5) * 0.5) / 3)
    # step 1: critic only; step 2: actor + Polyak targets
    actor0 = [p.clone() for p in L.actor.parameters()]
    tgt0 = [p.clone() for p in L.target_critic.parameters()]
    al, cl = L.update(s, a, r, s2, d)
    assert al is None and torch.isfinite(cl)
    assert all(torch.equal(p, q) for p, q in zip(L.actor.parameters(), actor0))
    assert all(torch.equal(p, q) for p, q in zip(L.target_critic.parameters(), tgt0))
    crit1 = [p.clone() for p in L.critic.parameters()]
    al, cl = L.update(s, a, r, s2, d)
    assert al is not None and torch.isfinite(al)
    assert any(not torch.equal(p, q) for p, q in zip(L.actor.parameters(), actor0))
    # Polyak: target = (1 - tau) * old_target + tau * current
    for tp, t0, p in zip(L.target_critic.parameters(), tgt0, L.critic.parameters()):
        assert torch.allclose(tp, (1 - cfg.tau_critic) * t0 + cfg.tau_critic * p, atol=1e-7)
    # repeated updates on one batch drive the critic loss down
    first = cl.item()
    for _ in range(200):
        al, cl = L.update(s, a, r, s2, d)
    assert cl.item() < 0.5 * first


def test_prioritized_buffer_semantics():
    from hockey_env_b200.td3 import DevicePrioritizedReplayBuffer, DeviceTD3Learner, TD3Config
    buf = DevicePrioritizedReplayBuffer(1000, device="cpu", seed=3)
    s, a, r, s2, d = _batch(300)
    buf.push(s, a, r, s2, d)
    assert len(buf) == 300 and torch.all(buf.weights[:300] == 1e8)           # new entries: the initial (maximum) weight
    o, *_ = buf.sample(128)
    assert o.shape == (128, 18) and buf.last_batch_inds.shape == (128,)
    pr = torch.rand(128) + 0.1
    inds = buf.last_batch_inds.clone()
    buf.update_priorities(pr)
    assert buf.last_batch_inds is None and torch.allclose(buf.weights[inds], pr[[int((inds == i).nonzero()[-1]) for i in inds]])
    buf.weights[:300] = 1e-3
    buf.weights[7] = 1.0                                                     # one dominant transition
    buf.sample(2000)
    share = (buf.last_batch_inds == 7).float().mean().item()
    assert abs(share - 1.0 / (1.0 + 299e-3)) < 0.05                           # sampled proportionally to its weight
    p = buf.get_last_probs()
    assert p.sum().item() == pytest.approx(1.0, rel=1e-5)
    s3, a3, r3, s4, d3 = _batch(10, seed=9)
    buf.push(s3, a3, r3, s4, d3)
    assert torch.all(buf.weights[300:310] == 1.0)                            # a new transition gets the current maximum
    # importance weights: (1 / (N p))^beta normalised by their maximum; priorities = clamped mean |TD| of the two heads
    L = DeviceTD3Learner(config=TD3Config(prioritized_replay=True), replay_buffer=buf, device="cpu")
    batch = buf.sample(64)
    probs = buf.get_last_probs()
    w = L.importance_weights()
    ref = (1.0 / (probs * buf.size)) ** 0.15
    assert torch.allclose(w, ref / ref.max())
    inds = buf.last_batch_inds.clone()
    L.update(*batch)
    assert buf.last_batch_inds is None
    assert torch.all(buf.weights[inds] >= 1e-6) and torch.all(buf.weights[inds] <= 1e6)
    # wrap-around of the ring keeps weights aligned with the slots they belong to
    big = DevicePrioritizedReplayBuffer(100, device="cpu")
    for k in range(5):
        big.push(*_batch(30, seed=k))
    assert len(big) == 100 and big.pos == 50


def test_checkpoint_format_matches_reference():
    """state_dict() has the reference's four entries and parameter names (rl/td3/agent.py:269-275, networks.py:36-70)."""
    from hockey_env_b200.td3 import DeviceTD3Learner
    L = DeviceTD3Learner(device="cpu")
    sd = L.state_dict()
    assert set(sd) == {"policy", "critic", "target_policy", "target_critic"}
    assert set(sd["policy"]) == {"fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias"}
    assert {"action_low", "action_high", "action_range", "q1.fc1.weight", "q2.fc3.bias"} <= set(sd["critic"])
    assert sd["critic"]["q1.fc1.weight"].shape == (256, 22) and sd["critic"]["q1.fc3.weight"].shape == (1, 256)
    L2 = DeviceTD3Learner(device="cpu")
    L2.load_state_dict(sd)
    x = torch.randn(5, 18)
    assert torch.equal(L.actor(x), L2.actor(x))


@pytest.mark.gpu
def test_td3_training_loop_on_the_env():
    """collect -> device replay (uniform and prioritized) -> learner, all on the GPU, starting from the reference's
    trained actor: the plumbing runs, losses stay finite, the buffers fill, parameters move."""
    import os
    import hockey_env_b200 as hk
    from hockey_env_b200 import td3
    npz = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "td3_actors.npz")
    for prio in (False, True):
        env = hk.HockeyVecEnv(512, device="cuda:0", seed=2, p2="weak")
        cfg = td3.TD3Config(prioritized_replay=prio, batch_size=256)
        Buf = td3.DevicePrioritizedReplayBuffer if prio else hk.DeviceReplayBuffer
        buf = Buf(512 * 40, device="cuda:0", seed=1)
        actor = hk.load_td3_actor(npz, device="cuda:0", name="stage_3").train()
        L = td3.DeviceTD3Learner(actor=actor, config=cfg, replay_buffer=buf, device="cuda:0")
        before = [p.clone() for p in L.actor.parameters()]
        loss = td3.train(env, L, ticks=60, updates_per_tick=2)
        assert torch.isfinite(loss) and len(buf) == 512 * 40 and L.train_step == 2 * (60 - 8)
        assert any(not torch.equal(p, q) for p, q in zip(L.actor.parameters(), before))
        assert env.stats()["env_steps"] == 512 * 60
        if prio:
            assert (buf.weights[:len(buf)] < 1e8).any()          # priorities were written back
        env.close()


@pytest.mark.gpu
def test_gym_vector_env_api():
    """HockeyGymVectorEnv: seeds (int / list), reset_mask, final_obs / final_info with their masks."""
    import hockey_env_b200 as hk
    n = 128
    v = hk.HockeyGymVectorEnv(n, mode=hk.Mode.TRAIN_SHOOTING, opponent="weak", seed=0)
    o1, i1 = v.reset(seed=11)
    o2, _ = v.reset(seed=11)
    o3, _ = v.reset(seed=[11 + k for k in range(n)])
    o4, _ = v.reset(seed=12)
    assert np.array_equal(o1[:, :16], o2[:, :16]) and np.array_equal(o1[:, :16], o3[:, :16]) and not np.array_equal(o1[:, :16], o4[:, :16])
    assert np.array_equal(o4[:-1, :16], o1[1:, :16])              # env i of seed 12 == env i+1 of seed 11
    assert set(i1) >= {"winner", "reward_closeness_to_puck", "reward_touch_puck", "reward_puck_direction"}
    mask = np.zeros(n, bool)
    mask[::3] = True
    for _ in range(5):
        v.step(np.zeros((n, 4), np.float32))
    o5, _ = v.reset(options={"reset_mask": mask})
    assert np.allclose(o5[mask, 0], -3.0) and np.all(o5[mask, 3:6] == 0)
    seen = 0
    for t in range(90):
        obs, rew, term, trunc, infos = v.step(np.zeros((n, 4), np.float32))
        assert np.array_equal(infos["_final_obs"], term) and np.array_equal(infos["_final_info"], term)
        if term.any():
            seen += int(term.sum())
            fi = infos["final_info"]
            assert set(fi) == {"winner", "reward_closeness_to_puck", "reward_touch_puck", "reward_puck_direction"}
            won = fi["winner"][term] != 0
            assert np.all(np.abs(rew[term][won]) > 9.0)           # +-10 on the tick a goal ends the episode
            # terminal obs vs first obs of the new episode (player 1 never moves under zero actions; the puck is re-drawn)
            assert np.all(np.any(infos["final_obs"][term, 12:14] != obs[term, 12:14], axis=1))
    assert seen >= n and not trunc.any()
    v.close()
    try:
        import gymnasium
        assert isinstance(v, gymnasium.vector.VectorEnv)
    except ImportError:
        pass

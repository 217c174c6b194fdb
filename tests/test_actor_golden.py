"""The reference's own trained TD3 actors as a fidelity probe of the whole step path.

The reference records, for each of its checkpoints, the win rate and mean return against the weak and the strong
BasicOpponent on the real pybox2d engine (Evaluator protocol, rl/utils/evaluator.py:10-35; numbers in
pretrained/stage_3/metrics/metrics.json and runs/20260216_113850_single_dual_eval_abcdefg_3(1)/metrics/metrics.json,
extracted into tests/golden/td3_actors.{npz,json} by tests/golden/make_actor_fixtures.py).  A policy trained on the real
engine only reaches those numbers on an engine that behaves like it -- contacts, keep/shoot, TOI bounces, rewards and
the opponent controller all enter.  Here the same actors play >= 10k COMPLETE episodes per pairing

  * on the CPU oracle (CPU tier: pins the oracle to the reference beyond the notebook fixtures), and
  * on the CUDA path through the public API (`-m gpu`), whose episode outcomes must in addition equal the oracle's
    episode by episode on the same seeds.

Tolerance: a recorded value is ONE 100-episode evaluation (binomial sd s = sqrt(p(1-p)/100) ~ 0.01-0.03) and it is the
best of 60 such evaluations during training (upward selection bias), so the measured rate must lie in
[recorded - 3.5 s, recorded + 2 s].  The recorded mean return is the mean of the same 100 episodes (a loss costs ~20
of return, so it moves with the win rate): same interval with s_r = (measured per-episode std of the return) / 10.
"""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

_G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(_G, "td3_actors.json")))
NPZ = os.path.join(_G, "td3_actors.npz")
N_ENVS, QUOTA = 1024, 10  # 10,240 complete episodes per pairing


def _cpu_actor(name):
    z = np.load(NPZ)
    w = {k[len(name) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + ".")}

    @torch.no_grad()
    def act(obs):  # ActorNetwork.forward (rl/td3/networks.py:17-20) in float32
        x = torch.from_numpy(obs)
        x = torch.tanh(x @ w["fc1.weight"].T + w["fc1.bias"])
        x = torch.tanh(x @ w["fc2.weight"].T + w["fc2.bias"])
        return torch.tanh(x @ w["fc3.weight"].T + w["fc3.bias"]).numpy()
    return act


def quota_play_oracle(O, act, weak, n=N_ENVS, k=QUOTA, seed=0, threads=8, record=None):
    """Evaluator protocol on the oracle: every env plays exactly k complete episodes (alternating sides)."""
    ora = O.OracleBatch(n, mode=O.MODE_NORMAL, seed=seed, n_threads=threads)
    obs = ora.reset(one_starting=(np.arange(n) % 2 == 0).astype(np.int8))
    count = np.zeros(n, np.int64)
    ret = np.zeros(n)
    wins = draws = losses = 0
    sum_ret = sum_ret2 = 0.0
    ticks = 0
    while (count < k).any():
        ro = ora.step(act(obs), O.POL_EXTERNAL, O.POL_WEAK if weak else O.POL_STRONG, O.STEP_AUTORESET)
        obs = ro["obs"]
        ticks += 1
        ret += ro["reward"]
        done = ro["done"].astype(bool)
        fin = done & (count < k)
        w = ro["info"][:, 0]
        wins += int((fin & (w == 1)).sum())
        draws += int((fin & (w == 0)).sum())
        losses += int((fin & (w == -1)).sum())
        sum_ret += float(ret[fin].sum())
        sum_ret2 += float((ret[fin] ** 2).sum())
        if record is not None:
            for i in np.nonzero(fin)[0]:
                record.append((ticks, int(i), int(w[i])))
        count += fin
        ret[done] = 0.0
        assert ticks < 252 * k + 1
    m = int(count.sum())
    return {"episodes": m, "win_rate": wins / m, "draw_rate": draws / m, "loss_rate": losses / m, "mean_return": sum_ret / m,
            "std_return": max(sum_ret2 / m - (sum_ret / m) ** 2, 0.0) ** 0.5, "ticks": ticks}


def check_against_record(res, name, weak):
    rec = META[name]["best_eval"]
    p_rec = rec["winrate_weak" if weak else "winrate_strong"]
    r_rec = rec["reward_weak" if weak else "reward_strong"]
    p = res["win_rate"]
    s = max(np.sqrt(p * (1 - p) / 100.0), 0.01)
    assert res["episodes"] >= 10_000
    assert p_rec - 3.5 * s <= p <= p_rec + 2.0 * s, (name, weak, p, p_rec, s)
    sr = max(res["std_return"] / 10.0, 0.05)
    assert r_rec - 3.5 * sr <= res["mean_return"] <= r_rec + 2.0 * sr, (name, weak, res["mean_return"], r_rec, sr)


@pytest.mark.parametrize("name", ["stage_3", "competition"])
@pytest.mark.parametrize("weak", [False, True])
def test_oracle_reaches_recorded_winrates(oracle, name, weak):
    res = quota_play_oracle(oracle, _cpu_actor(name), weak)
    print(name, "weak" if weak else "strong", res)
    check_against_record(res, name, weak)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["stage_3", "competition"])
@pytest.mark.parametrize("weak", [False, True])
def test_cuda_reaches_recorded_winrates(name, weak):
    """hk.evaluate (the Evaluator protocol over HockeyVecEnv) with load_td3_actor on the fixture weights."""
    import hockey_env_b200 as hk
    actor = hk.load_td3_actor(NPZ, device="cuda:0", name=name)
    res = hk.evaluate(actor, n_episodes=N_ENVS * QUOTA, opponent="weak" if weak else "strong", num_envs=N_ENVS, seed=0)
    print(name, "weak" if weak else "strong", res)
    assert res["episodes"] == N_ENVS * QUOTA and res["episodes_per_env"] == QUOTA
    check_against_record(res, name, weak)


@pytest.mark.gpu
def test_cuda_and_oracle_play_the_same_games(oracle):
    """Same seeds, same actor (evaluated once, on the GPU, and fed to both engines): the CUDA env and the oracle must
    produce the same observations every tick and therefore the same episode outcomes."""
    import hockey_env_b200 as hk
    O = oracle
    n = 512
    actor = hk.load_td3_actor(NPZ, device="cuda:0", name="stage_3")
    env = hk.HockeyVecEnv(n, device="cuda:0", seed=11, p2="strong")
    ora = O.OracleBatch(n, mode=O.MODE_NORMAL, seed=11, n_threads=8)
    side = (np.arange(n) % 2 == 0).astype(np.int8)
    obs, _ = env.reset(one_starting=torch.from_numpy(side).cuda())
    assert np.array_equal(obs.cpu().numpy(), ora.reset(one_starting=side))
    wins = 0
    for t in range(400):
        with torch.no_grad():
            a = actor(obs).contiguous()
        obs, reward, done, _, info = env.step(a)
        ro = ora.step(a.cpu().numpy(), O.POL_EXTERNAL, O.POL_STRONG, O.STEP_AUTORESET)
        assert np.array_equal(obs.cpu().numpy(), ro["obs"]), f"obs differs at tick {t}"
        assert np.array_equal(done.cpu().numpy(), ro["done"]), f"done differs at tick {t}"
        assert np.array_equal(info["winner"].cpu().numpy(), ro["info"][:, 0].astype(np.float32)), f"winner differs at tick {t}"
        assert np.array_equal(reward.cpu().numpy(), ro["reward"].astype(np.float32)), f"reward differs at tick {t}"
        wins += int((info["winner"] == 1).sum().item())
    assert wins > 1000  # ~5 episodes per env, ~90 % won


@pytest.mark.gpu
def test_model_evaluator_protocol():
    """ModelEvaluator defaults (model_evaluation/model_evaluator.py:234-235): 300 episodes, seed 123, both opponents."""
    import hockey_env_b200 as hk
    actor = hk.load_td3_actor(NPZ, device="cuda:0", name="competition")
    r = hk.evaluate_model(actor)
    assert r["episodes"] == 300
    assert r["wr_strong"] > 0.85 and r["wr_weak"] > 0.9 and r["ret_strong"] > 6.5 and r["ret_weak"] > 8.0


def test_load_td3_actor_from_fixture_cpu():
    """load_td3_actor maps the fixture (and a TD3Agent.save-style checkpoint) onto the reference architecture."""
    from hockey_env_b200.actor import load_td3_actor
    a = load_td3_actor(NPZ, device="cpu", name="stage_3")
    x = torch.from_numpy(np.random.default_rng(0).normal(size=(7, 18)).astype(np.float32))
    with torch.no_grad():
        y = a(x).numpy()
    assert y.shape == (7, 4) and np.abs(y).max() <= 1.0
    assert np.allclose(y, _cpu_actor("stage_3")(x.numpy()), atol=1e-6)
    with pytest.raises(ValueError):
        load_td3_actor(NPZ, device="cpu")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "td3_best.pt")
        torch.save({"policy": a.state_dict(), "critic": {}}, p)
        b = load_td3_actor(p, device="cpu")
        with torch.no_grad():
            assert np.array_equal(b(x).numpy(), y)

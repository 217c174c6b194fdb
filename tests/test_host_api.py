"""CPU tier: the C-ABI library loads and exports every symbol include/hockey_b200.h declares (no compute calls
without a GPU), the host-side mirror keeps the reference's names / errors, and the package has no CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import hockey_env_b200 as hk
    header = open(os.path.join(ROOT, "include", "hockey_b200.h")).read()
    declared = set(re.findall(r"\b(hk_[a-z_0-9]+)\s*\(", header))
    assert declared == set(hk._lib.EXPORTS), declared ^ set(hk._lib.EXPORTS)
    L = hk.load_library()
    for sym in declared:
        assert getattr(L, sym) is not None
    assert b"sm_100a" in L.hk_version()


def test_error_convention_without_gpu():
    import torch
    import hockey_env_b200 as hk
    L = hk.load_library()
    h = ctypes.c_void_p()
    assert L.hk_create(ctypes.byref(h), 0, 0, 1, 0, 0, 0) == hk._lib.HK_E_INVALID      # n_envs <= 0
    assert b"n_envs" in L.hk_last_error()
    assert L.hk_create(ctypes.byref(h), 4, 7, 1, 0, 0, 0) == hk._lib.HK_E_INVALID      # bad mode
    assert b"not a valid value for Mode" in L.hk_last_error()
    if not torch.cuda.is_available():
        assert L.hk_create(ctypes.byref(h), 4, 0, 1, 0, 0, 0) == hk._lib.HK_E_NODEVICE
        with pytest.raises(hk.HockeyLibraryError):
            hk.HockeyVecEnv(4)                                                         # loud, no CPU fallback
        with pytest.raises(hk.HockeyLibraryError):
            hk.HockeyEnv()


def test_state_record_layout_constants_agree():
    import hockey_env_b200 as hk
    import oracle_lib as O
    header = open(os.path.join(ROOT, "include", "hockey_b200.h")).read()
    m = re.search(r"HK_STATE_WORDS\s*=\s*64\s*\+\s*27\s*\*\s*8", header)
    assert m and hk._lib.STATE_WORDS == O.STATE_WORDS == 64 + 27 * 8


def test_mode_parsing_matches_reference():
    """mode setter semantics of hockey_env.py:758-779."""
    from hockey_env_b200.env import Mode, _as_mode
    assert _as_mode(Mode.TRAIN_DEFENSE) is Mode.TRAIN_DEFENSE
    assert _as_mode("TRAIN_SHOOTING") is Mode.TRAIN_SHOOTING
    assert _as_mode(0) is Mode.NORMAL
    with pytest.raises(ValueError, match="is not a valid name for Mode"):
        _as_mode("nope")
    with pytest.raises(ValueError, match="is not a valid value for Mode"):
        _as_mode(9)
    with pytest.raises(TypeError):
        _as_mode(1.0)


def test_discrete_to_continuous_action():
    """hockey_env.py:637-656."""
    from hockey_env_b200 import HockeyVecEnv
    f = HockeyVecEnv.discrete_to_continous_action
    assert f(0) == [0.0, 0.0, 0.0, 0.0]
    assert f(1) == [-1.0, 0.0, 0.0, 0.0] and f(2) == [1.0, 0.0, 0.0, 0.0]
    assert f(3) == [0.0, -1.0, 0.0, 0.0] and f(4) == [0.0, 1.0, 0.0, 0.0]
    assert f(5) == [0.0, 0.0, -1.0, 0.0] and f(6) == [0.0, 0.0, 1.0, 0.0]
    assert f(7) == [0.0, 0.0, 0.0, 1.0]
    assert f(7, keep_mode=False) == [0.0, 0.0, 0.0]


def test_basic_opponent_numpy_matches_oracle_controller(oracle):
    """The host-side BasicOpponent.act (reference semantics, hockey_env.py:787-833) and the in-kernel / oracle
    controller give the same action for the same observation and phase."""
    import hockey_env_b200 as hk
    O = oracle
    b = O.OracleBatch(8, mode=0, seed=3)
    for _ in range(40):
        b.step(None, O.POL_STRONG, O.POL_WEAK, O.STEP_AUTORESET)
    obs, obs2 = b.get_obs()
    s0 = b.get_state()
    phases = s0[:, O.S_PHASE:O.S_PHASE + 4].copy().view(np.float64)          # [n, 2]
    out = b.step(None, O.POL_STRONG, O.POL_WEAK, 0)
    phases_after = b.get_state()[:, O.S_PHASE:O.S_PHASE + 4].copy().view(np.float64)
    # replay player 1's controller on the host with the same phase increment
    for i in range(8):
        opp = hk.BasicOpponent(weak=False)
        opp.phase = phases[i, 0]
        inc = phases_after[i, 0] - phases[i, 0]
        state = np.random.get_state()
        try:
            class _Fixed:
                pass
            orig = np.random.uniform
            np.random.uniform = lambda lo, hi: inc
            a = opp.act(obs[i].astype(np.float64))
        finally:
            np.random.uniform = orig
            np.random.set_state(state)
        assert a.shape == (4,) and np.all(np.abs(a[:3]) <= 1)
        # the oracle applied exactly this action: reproduce the tick from the saved state with it as external input
    b2 = O.OracleBatch(8, mode=0, seed=3)
    b2.set_state(s0)
    acts = np.zeros((8, 8), np.float32)
    for i in range(8):
        for k, (ob, weak) in enumerate(((obs[i], False), (obs2[i], True))):
            opp = hk.BasicOpponent(weak=weak)
            opp.phase = phases[i, k]
            inc = phases_after[i, k] - phases[i, k]
            orig = np.random.uniform
            np.random.uniform = lambda lo, hi: inc
            try:
                acts[i, 4 * k:4 * k + 4] = opp.act(ob.astype(np.float64))
            finally:
                np.random.uniform = orig
    out2 = b2.step(acts, O.POL_EXTERNAL, O.POL_EXTERNAL, 0)
    assert np.array_equal(out["obs"], out2["obs"])


def test_registry_mirrors_the_reference_ids():
    """hockey_env.py:889-903 registers 'Hockey-v0' (HockeyEnv, mode 0) and 'Hockey-One-v0' (HockeyEnv_BasicOpponent,
    mode 0, strong opponent); make()/spec() resolve the same ids with the same defaults."""
    import hockey_env_b200 as hk
    cls, kw = hk.spec("Hockey-v0")
    assert cls is hk.HockeyEnv and kw == {"mode": 0}
    cls, kw = hk.spec("Hockey-One-v0")
    assert cls is hk.HockeyEnv_BasicOpponent and kw == {"mode": 0, "weak_opponent": False}
    with pytest.raises(ValueError):
        hk.spec("Hockey-Two-v0")
    import torch
    if not torch.cuda.is_available():  # no CPU fallback: construction must fail loudly, not silently
        with pytest.raises(hk.HockeyLibraryError):
            hk.make("Hockey-v0")

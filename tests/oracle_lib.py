"""ctypes binding of the CPU oracle (oracle/libhockey_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg.  The product package (hockey_env_b200) never imports this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ORACLE_DIR, "libhockey_oracle.so")

OBS_DIM, STATE_WORDS, STATS_DIM, N_PAIRS, CONTACT_WORDS = 18, 64 + 27 * 8, 16, 27, 8
S_R1, S_R2, S_PUCK, S_SLEEP, S_FLAGS, S_TIME, S_HAS1, S_HAS2, S_PFORCE, S_FAT, S_MOVED, S_PHASE, S_EPISODE, S_TICK, S_RET, S_CONTACT = (
    0, 8, 16, 22, 25, 26, 27, 28, 29, 31, 43, 44, 48, 49, 50, 64)
MODE_NORMAL, MODE_TRAIN_SHOOTING, MODE_TRAIN_DEFENSE = 0, 1, 2
POL_EXTERNAL, POL_WEAK, POL_STRONG, POL_RANDOM, POL_ZERO = 0, 1, 2, 3, 4
STEP_AUTORESET = 1


def build(force=False):
    srcs = [os.path.join(_ORACLE_DIR, f) for f in ("b2mini.cpp", "b2mini.h", "hockey_oracle.cpp", "Makefile")]
    srcs.append(os.path.join(_ROOT, "include", "hockey_b200.h"))
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs if os.path.exists(s))):
        return _SO
    if not os.path.exists(os.path.join(_ORACLE_DIR, "hockey_oracle.cpp")):
        return _SO
    subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
        L.hko_create.restype = vp
        L.hko_create.argtypes = [i64, i32, i32, u64, i64, i32]
        L.hko_destroy.argtypes = [vp]
        L.hko_set_modes.argtypes = [i32, i32]
        L.hko_reset.argtypes = [vp, vp, vp, vp]
        L.hko_reset_seeded.argtypes = [vp, vp, vp, vp, vp]
        L.hko_reset_with_draws.argtypes = [vp, i64, i32, vp]
        L.hko_get_obs.argtypes = [vp, vp, vp]
        L.hko_step.argtypes = [vp, vp, i32, i32, i32, i32] + [vp] * 8
        L.hko_rollout.argtypes = [vp, i32, i32, i32]
        L.hko_get_state.argtypes = [vp, vp]
        L.hko_set_state.argtypes = [vp, vp]
        L.hko_set_obs_state.argtypes = [vp, vp]
        L.hko_get_stats.argtypes = [vp, vp]
        L.hko_clear_stats.argtypes = [vp]
        L.hko_scene_constants.argtypes = [vp, vp]
        L.hko_scene_polygon.argtypes = [vp, i32, vp]
        L.hko_scene_polygon.restype = i32
        L.hko_sincosf.argtypes = [vp, i64, vp, vp]
        L.hko_trig_check.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """A batch of oracle envs with the same calling shape as the CUDA library."""

    def __init__(self, n, mode=MODE_NORMAL, keep_mode=True, seed=0, env_id_offset=0, n_threads=1):
        self.n = int(n)
        self.L = lib()
        self.h = self.L.hko_create(self.n, int(mode), int(bool(keep_mode)), int(seed), int(env_id_offset), int(n_threads))

    def __del__(self):
        try:
            if self.h:
                self.L.hko_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self, mask=None, one_starting=None, seeds=None):
        obs = np.zeros((self.n, OBS_DIM), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        o = None if one_starting is None else np.ascontiguousarray(one_starting, np.int8)
        sd = None if seeds is None else np.ascontiguousarray(seeds, np.int64)
        self.L.hko_reset_seeded(self.h, _p(m), _p(o), _p(sd), _p(obs))
        return obs

    def reset_with_draws(self, index, draws, one_starting=-1):
        d = np.ascontiguousarray(draws, np.float64)
        self.L.hko_reset_with_draws(self.h, int(index), int(one_starting), _p(d))

    def get_obs(self):
        obs = np.zeros((self.n, OBS_DIM), np.float32)
        obs2 = np.zeros((self.n, OBS_DIM), np.float32)
        self.L.hko_get_obs(self.h, _p(obs), _p(obs2))
        return obs, obs2

    def step(self, action=None, p1=POL_EXTERNAL, p2=POL_EXTERNAL, flags=0):
        n = self.n
        a, stride = None, 0
        if action is not None:
            a = np.ascontiguousarray(action, np.float32)
            stride = a.shape[1]
        out = dict(obs=np.zeros((n, 18), np.float32), obs2=np.zeros((n, 18), np.float32),
                   reward=np.zeros(n, np.float64), reward2=np.zeros(n, np.float64), done=np.zeros(n, np.uint8),
                   info=np.zeros((n, 4), np.float64), info2=np.zeros((n, 4), np.float64),
                   final_obs=np.zeros((n, 18), np.float32))
        self.L.hko_step(self.h, _p(a), stride, p1, p2, flags, _p(out["obs"]), _p(out["obs2"]), _p(out["reward"]),
                        _p(out["reward2"]), _p(out["done"]), _p(out["info"]), _p(out["info2"]), _p(out["final_obs"]))
        return out

    def rollout(self, k, p1, p2):
        self.L.hko_rollout(self.h, int(k), int(p1), int(p2))

    def get_state(self):
        s = np.zeros((self.n, STATE_WORDS), np.uint32)
        self.L.hko_get_state(self.h, _p(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.uint32)
        assert s.shape == (self.n, STATE_WORDS)
        self.L.hko_set_state(self.h, _p(s))

    def set_obs_state(self, obs18):
        s = np.ascontiguousarray(obs18, np.float64)
        self.L.hko_set_obs_state(self.h, _p(s))

    def stats(self):
        s = np.zeros(STATS_DIM, np.float64)
        self.L.hko_get_stats(self.h, _p(s))
        return s

    def clear_stats(self):
        self.L.hko_clear_stats(self.h)

    def scene_constants(self):
        s = np.zeros(16, np.float32)
        self.L.hko_scene_constants(self.h, _p(s))
        return s

    def scene_polygon(self, f):
        buf = np.zeros(80, np.float32)
        n = self.L.hko_scene_polygon(self.h, int(f), _p(buf))
        return dict(verts=buf[:2 * n].reshape(n, 2).copy(), normals=buf[2 * n:4 * n].reshape(n, 2).copy(),
                    centroid=buf[4 * n:4 * n + 2].copy(), pos=buf[4 * n + 2:4 * n + 4].copy(),
                    fat=buf[4 * n + 4:4 * n + 8].copy())


def set_modes(trig_mode=0, static_drift=0):
    lib().hko_set_modes(int(trig_mode), int(static_drift))

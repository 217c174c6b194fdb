"""ctypes binding of tests/hostsim (the product's device headers compiled for the host).

TEST INFRASTRUCTURE: lets the CPU test tier compare the kernel arithmetic with the oracle bit-for-bit.
Never imported by the hockey_env_b200 package.
"""
import ctypes as C
import os
import subprocess
import numpy as np
import oracle_lib as O

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DIR = os.path.join(_ROOT, "tests", "hostsim")
_SO = os.path.join(_DIR, "libhostsim.so")
_CSRC = os.path.join(_ROOT, "hockey_env_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(_DIR, "hostsim.cpp"), os.path.join(_ROOT, "include", "hockey_b200.h")]
    srcs += [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                           "-x", "c++", "-o", _SO, os.path.join(_DIR, "hostsim.cpp")])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i64, u64, i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int
        L.hs_create.restype = vp
        L.hs_create.argtypes = [i64, i32, i32, u64, i64]
        L.hs_destroy.argtypes = [vp]
        L.hs_reset.argtypes = [vp, vp, vp, vp]
        L.hs_reset_seeded.argtypes = [vp, vp, vp, vp, vp]
        L.hs_get_obs.argtypes = [vp, vp, vp]
        L.hs_step.argtypes = [vp, vp, i32, i32, i32, i32] + [vp] * 8
        L.hs_get_state.argtypes = [vp, vp]
        L.hs_set_state.argtypes = [vp, vp]
        L.hs_set_obs_state.argtypes = [vp, vp]
        L.hs_get_stats.argtypes = [vp, vp]
        L.hs_clear_stats.argtypes = [vp]
        L.hs_scene.argtypes = [vp, vp, i64]
        L.hs_scene_size.restype = i64
        L.hs_set_fast.argtypes = [vp, i32]
        L.hs_fast_counts.argtypes = [vp, vp]
        L.hs_set_mid_budget.argtypes = [vp, i32]
        _lib = L
    return _lib


_p = O._p


class HostSimBatch:
    def __init__(self, n, mode=0, keep_mode=True, seed=0, env_id_offset=0, fast=False):
        self.n = int(n)
        self.L = lib()
        self.h = self.L.hs_create(self.n, int(mode), int(bool(keep_mode)), int(seed), int(env_id_offset))
        self.L.hs_set_fast(self.h, int(bool(fast)))

    def __del__(self):
        try:
            if self.h:
                self.L.hs_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def reset(self, mask=None, one_starting=None, seeds=None):
        obs = np.zeros((self.n, 18), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        o = None if one_starting is None else np.ascontiguousarray(one_starting, np.int8)
        sd = None if seeds is None else np.ascontiguousarray(seeds, np.int64)
        self.L.hs_reset_seeded(self.h, _p(m), _p(o), _p(sd), _p(obs))
        return obs

    def get_obs(self):
        obs = np.zeros((self.n, 18), np.float32)
        obs2 = np.zeros((self.n, 18), np.float32)
        self.L.hs_get_obs(self.h, _p(obs), _p(obs2))
        return obs, obs2

    def step(self, action=None, p1=0, p2=0, flags=0):
        n = self.n
        a, stride = None, 0
        if action is not None:
            a = np.ascontiguousarray(action, np.float32)
            stride = a.shape[1]
        out = dict(obs=np.zeros((n, 18), np.float32), obs2=np.zeros((n, 18), np.float32),
                   reward=np.zeros(n, np.float32), reward2=np.zeros(n, np.float32), done=np.zeros(n, np.uint8),
                   info=np.zeros((n, 4), np.float32), info2=np.zeros((n, 4), np.float32),
                   final_obs=np.zeros((n, 18), np.float32))
        self.L.hs_step(self.h, _p(a), stride, p1, p2, flags, _p(out["obs"]), _p(out["obs2"]), _p(out["reward"]),
                       _p(out["reward2"]), _p(out["done"]), _p(out["info"]), _p(out["info2"]), _p(out["final_obs"]))
        return out

    def get_state(self):
        s = np.zeros((self.n, O.STATE_WORDS), np.uint32)
        self.L.hs_get_state(self.h, _p(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.uint32)
        self.L.hs_set_state(self.h, _p(s))

    def set_obs_state(self, obs18):
        s = np.ascontiguousarray(obs18, np.float32)
        self.L.hs_set_obs_state(self.h, _p(s))

    def fast_counts(self):
        c = np.zeros(4, np.int64)
        self.L.hs_fast_counts(self.h, _p(c))
        return int(c[0]), int(c[1]), int(c[2])

    def touch_count(self):
        c = np.zeros(4, np.int64)
        self.L.hs_fast_counts(self.h, _p(c))
        return int(c[3])

    def stats(self):
        s = np.zeros(16, np.float64)
        self.L.hs_get_stats(self.h, _p(s))
        return s

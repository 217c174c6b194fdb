"""Helpers shared by the parity tests: comparison of canonical state records."""
import numpy as np
import oracle_lib as O

# words of the canonical record that hold float32 values (compared numerically so that -0.0 == +0.0:
# the kernel's early solver exit can leave a zero impulse with the other sign bit -- numerically
# identical, and no operation on this path depends on the sign of a zero; see DESIGN.md "Numerics")
_FLOAT_WORDS = np.zeros(O.STATE_WORDS, bool)
_FLOAT_WORDS[0:25] = True
_FLOAT_WORDS[29:43] = True
_FLOAT_WORDS[54:56] = True
for _p in range(O.N_PAIRS):
    _FLOAT_WORDS[O.S_CONTACT + 8 * _p + 4:O.S_CONTACT + 8 * _p + 8] = True


def state_mismatches(a, b):
    """indices [env, word] where two [n, STATE_WORDS] uint32 records differ (floats: numeric equality)."""
    a = np.ascontiguousarray(a, np.uint32)
    b = np.ascontiguousarray(b, np.uint32)
    neq_bits = a != b
    fa, fb = a.view(np.float32), b.view(np.float32)
    neq_num = ~((fa == fb) | (np.isnan(fa) & np.isnan(fb)))
    neq = np.where(_FLOAT_WORDS[None, :], neq_num, neq_bits)
    return np.argwhere(neq)


def outputs_equal(ro, rh):
    """oracle step dict (float64 rewards/info) vs kernel-side dict (float32): exact after the f32 cast."""
    bad = []
    for k in ("obs", "obs2", "done", "final_obs"):
        if not np.array_equal(ro[k], rh[k]):
            bad.append(k)
    for k in ("reward", "reward2", "info", "info2"):
        if not np.array_equal(ro[k].astype(np.float32), rh[k]):
            bad.append(k)
    return bad

"""Build the CUDA library (sm_100a only) in-tree: hockey_env_b200/libhockey_b200.so.

csrc/hk_lib.cu is compiled twice and the two objects are linked into one shared library (see the comment at
`namespace hkinl` in hk_lib.cu): pass 1 is the whole library with the large device helpers kept as functions, pass 2
(-DHK_TU_INLINE -DHK_INLINE_ALL) holds only the general-tier and touch-tier kernels with every helper inlined, and is built
twice (register budgets for blocks of up to 384 and up to 256 threads).  The three compilations run in parallel.
"""
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
SO = os.path.join(_PKG, "libhockey_b200.so")
OBJ_DIR = os.path.join(_ROOT, "build")

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # no FMA contraction: bit-comparable with an SSE build of Box2D and with the oracle
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "--cudart", "shared"]  # libcudart.so of the process (torch ships one): no second copy of the runtime
# inlining groups (hk_math.cuh), measured on the B200 (profiles/README.md, "inlining"): pass 1 keeps the helpers as
# functions except force shaping / info / tick epilogue, the sin-cos polynomial and the swept-AABB update (k_fast -15 %;
# inlining its contact-list walk makes it 2.2x SLOWER); pass 2 inlines everything except the contact-list walk of Collide
PASS1_FLAGS = ["-DHK_IN_FASTA", "-DHK_IN_MATH", "-DHK_IN_FASTW2"]
PASS2_FLAGS = ["-DHK_TU_INLINE", "-DHK_INLINE_ALL", "-DHK_OUT_COLLIDE"]
PASS2B_FLAGS = ["-DHK_INL_NS=hkinl256", "-DHK_GENERAL_BOUND=256"]  # on top of pass 2's: the same kernel for blocks of <= 256 threads


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(_ROOT, "include", "hockey_b200.h"), os.path.abspath(__file__)]


def is_stale(so=SO):
    if not os.path.exists(so):
        return True
    t = os.path.getmtime(so)
    return any(os.path.getmtime(s) > t for s in sources())


def build_cuda(force=False, verbose=False, so=SO, pass2_flags=None, tag=""):
    """Compile csrc/hk_lib.cu (two passes, in parallel) with nvcc -- cross-compiles without a GPU. Returns the .so path."""
    if not force and not is_stale(so):
        return so
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build hockey_env_b200/libhockey_b200.so")
    os.makedirs(OBJ_DIR, exist_ok=True)
    src = os.path.join(CSRC, "hk_lib.cu")
    v = ["-Xptxas", "-v"] if verbose else []
    o1, o2, o3 = (os.path.join(OBJ_DIR, f"{n}{tag}.o") for n in ("hk_lib", "hk_inl", "hk_inl256"))
    p2 = PASS2_FLAGS if pass2_flags is None else list(pass2_flags)
    procs = [subprocess.Popen([nvcc] + NVCC_FLAGS + v + PASS1_FLAGS + ["-c", "-o", o1, src]),
             subprocess.Popen([nvcc] + NVCC_FLAGS + v + p2 + ["-c", "-o", o2, src]),
             subprocess.Popen([nvcc] + NVCC_FLAGS + v + p2 + PASS2B_FLAGS + ["-c", "-o", o3, src])]
    rcs = [p.wait() for p in procs]
    if any(rcs):
        raise RuntimeError(f"nvcc failed (pass 1 rc={rcs[0]}, pass 2 rc={rcs[1]}, pass 2b rc={rcs[2]})")
    subprocess.check_call([nvcc] + LINK_FLAGS + ["-o", so, o1, o2, o3])
    return so


if __name__ == "__main__":
    import sys
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))

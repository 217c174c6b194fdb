"""Build the CUDA library (sm_100a only) in-tree: hockey_env_b200/libhockey_b200.so."""
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
SO = os.path.join(_PKG, "libhockey_b200.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",  # no FMA contraction: bit-comparable with an SSE build of Box2D and with the oracle
    "-Xcompiler", "-fPIC", "-shared",
    "--cudart", "shared",  # libcudart.so of the process (torch ships one): no second copy of the runtime in this library
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(_ROOT, "include", "hockey_b200.h")]


def is_stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build_cuda(force=False, verbose=False):
    """Compile csrc/hk_lib.cu with nvcc (cross-compiles without a GPU). Returns the .so path."""
    if not force and not is_stale():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build hockey_env_b200/libhockey_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, os.path.join(CSRC, "hk_lib.cu")]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    import sys
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))

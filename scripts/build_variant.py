"""A/B builds of the library with different inlining groups: scripts/build_variant.py NAME [--p1 "flags"] [--p2 "flags"]
-> build_variants/lib_NAME.so (run them with scripts/ab_libs.sh).  --p1 / --p2: the inlining-group flags of pass 1 / pass 2 of
hk_lib.cu, REPLACING build.py's defaults (e.g. --p2="-DHK_OUT_TOI -DHK_OUT_COLLIDE", --p1="-DHK_IN_MATH"; "default" or
absent = build.py's set); see hk_math.cuh for the groups."""
import argparse
import os
import shlex
import shutil
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hockey_env_b200 import build as hb

ap = argparse.ArgumentParser()
ap.add_argument("name")
ap.add_argument("--p1", default="")
ap.add_argument("--p2", default="")
a = ap.parse_args()
nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
os.makedirs(hb.OBJ_DIR, exist_ok=True)
os.makedirs(os.path.join(hb._ROOT, "build_variants"), exist_ok=True)
src = os.path.join(hb.CSRC, "hk_lib.cu")
o1 = os.path.join(hb.OBJ_DIR, f"v_{a.name}_1.o")
o2 = os.path.join(hb.OBJ_DIR, f"v_{a.name}_2.o")
procs = []
base1 = os.path.join(hb.OBJ_DIR, "hk_lib.o")
if a.p1 or not os.path.exists(base1):
    procs.append(subprocess.Popen([nvcc] + hb.NVCC_FLAGS + (shlex.split(a.p1) if a.p1 != "default" else hb.PASS1_FLAGS) + ["-c", "-o", o1, src]))
else:
    o1 = base1  # the default pass 1 (python hockey_env_b200/build.py --force leaves it in build/)
base2 = os.path.join(hb.OBJ_DIR, "hk_inl.o")
if a.p2 or not os.path.exists(base2):
    procs.append(subprocess.Popen([nvcc] + hb.NVCC_FLAGS + (["-DHK_TU_INLINE", "-DHK_INLINE_ALL"] + shlex.split(a.p2) if a.p2 != "default" else hb.PASS2_FLAGS) + ["-c", "-o", o2, src]))
else:
    o2 = base2
o3 = os.path.join(hb.OBJ_DIR, f"v_{a.name}_3.o")
base3 = os.path.join(hb.OBJ_DIR, "hk_inl256.o")
if a.p2 or not os.path.exists(base3):
    procs.append(subprocess.Popen([nvcc] + hb.NVCC_FLAGS + (["-DHK_TU_INLINE", "-DHK_INLINE_ALL"] + shlex.split(a.p2) if a.p2 != "default" else hb.PASS2_FLAGS) +
                                  hb.PASS2B_FLAGS + ["-c", "-o", o3, src]))
else:
    o3 = base3
if any([p.wait() for p in procs]):
    sys.exit("nvcc failed")
so = os.path.join(hb._ROOT, "build_variants", f"lib_{a.name}.so")
subprocess.check_call([nvcc] + hb.LINK_FLAGS + ["-o", so, o1, o2, o3])
print(so)

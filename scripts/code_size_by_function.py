"""Static code size of a kernel by source function (inlined code attributed to the function whose lines it came from).
The general tier is instruction-fetch bound (492 KB of SASS, 32 KB L1.5 instruction cache), so this is the map of what
to shrink or un-inline.  Usage: python scripts/code_size_by_function.py [kernel-substring] (needs cuobjdump / nvdisasm)."""
import bisect
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "hockey_env_b200", "libhockey_b200.so")
KERNEL = sys.argv[1] if len(sys.argv) > 1 else "k_generalILi1"
csrc = os.path.join(ROOT, "hockey_env_b200", "csrc")


def funcs(path):
    out = []
    for n, l in enumerate(open(path), 1):
        if l.startswith(("HK_HD", "__global__", "__device__", "template", "static")):
            m = re.search(r"(\w+)\(", l)
            if m:
                out.append((n, m.group(1)))
    return out


fmap = {f: funcs(os.path.join(csrc, f)) for f in os.listdir(csrc)}
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cur, fl, ln, counts, total = None, None, 0, {}, 0
for l in sass.splitlines():
    if l.startswith(".text."):
        cur = l
        continue
    if cur is None or KERNEL not in cur:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        fl, ln = os.path.basename(m.group(1)), int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        name = "?"
        if fl in fmap and fmap[fl]:
            starts = [x[0] for x in fmap[fl]]
            k = bisect.bisect_right(starts, ln) - 1
            if k >= 0:
                name = fmap[fl][k][1]
        counts[(fl, name)] = counts.get((fl, name), 0) + 1
        total += 1
print(f"{KERNEL}: {total} SASS instructions = {total * 16 / 1024:.0f} KB")
for (f, n), c in sorted(counts.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{c:7d} {100 * c / total:5.1f}%  {c * 16 / 1024:6.1f} KB  {f}:{n}")

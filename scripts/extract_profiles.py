"""Turns gpurun_out/<tag>_full.ncu-rep (+ launch list, bench JSONs) into the small tracked files under profiles/.
Run in the build container (ncu can read reports without a GPU):  python scripts/extract_profiles.py [tag]
(tag = r1_final (default), r1b, ...: the capture set of one gpurun evidence call)"""
import bisect
import csv
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1_final"
SHORT = TAG.replace("_final", "")  # bench_<short>_*.json
REP = os.path.join(OUT, TAG + "_full.ncu-rep")
csv.field_size_limit(10 ** 9)


def ncu(*args):
    return subprocess.run(["ncu", "-i", REP, *args], capture_output=True, text=True).stdout


def main():
    os.makedirs(PROF, exist_ok=True)
    pc = "phase_cycles.txt" if TAG == "r1_final" else "phase_cycles_%s.txt" % SHORT
    for src, dst in ((TAG + "_launches.csv", TAG + "_launches.csv"), ("bench_%s_final.json" % SHORT, SHORT + "_bench_1gpu.json"),
                     ("bench_%s_131k.json" % SHORT, SHORT + "_bench_1gpu_131k_envs.json"), ("bench_%s_32k.json" % SHORT, SHORT + "_bench_1gpu_32k_envs.json"),
                     ("bench_%s_1M.json" % SHORT, SHORT + "_bench_1gpu_1M_envs.json"), ("bench_%s_2gpu.json" % SHORT, SHORT + "_bench_2gpu.json"),
                     ("bench_%s_ref.json" % SHORT, SHORT + "_bench_reference_arm.json"), (pc, SHORT + "_phase_cycles.txt")):
        if os.path.exists(os.path.join(OUT, src)):
            shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))
    rows = list(csv.reader(ncu("--page", "raw", "--csv").splitlines()))
    hdr, units = rows[0], rows[1]
    pref = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct",
            "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__occupancy_limit", "smsp__average_warps_issue_stalled",
            "smsp__average_warp_latency", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate", "sm__throughput.avg.pct",
            "gpu__dram_throughput", "sm__inst_executed_pipe_fp64", "smsp__inst_executed_op_local", "sm__cycles_elapsed.max",
            "smsp__cycles_active.avg", "l1tex__t_bytes_pipe_lsu_mem_local")
    keep = [i for i, h in enumerate(hdr) if h in ("Block Size", "Grid Size") or h.startswith(pref)]
    with open(os.path.join(PROF, TAG + "_raw_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")][:44] for r in rows[2:]])
        for i in keep:
            w.writerow([hdr[i], units[i]] + [r[i] for r in rows[2:]])
    data, fname, h2, kern = {}, None, None, None
    for r in csv.reader(ncu("--page", "source", "--csv", "--print-source", "cuda,sass").splitlines()):
        if not r or r[0] == "Kernel Name":
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            kern = "k_fast" if "k_fast" in r[1] else "k_general"
        elif r[0] == "Line No":
            h2 = r
        elif h2 and len(r) >= 10 and r[0] != "":
            try:
                line, smp, ie, te = int(r[0]), float(r[6] or 0), float(r[7] or 0), float(r[8] or 0)
            except ValueError:
                continue
            a = data.setdefault(kern, {}).setdefault((fname, line), [0, 0, 0])
            a[0] += ie
            a[1] += te
            a[2] += smp

    def funcs(path):
        out = []
        for n, l in enumerate(open(path), 1):
            if l.startswith(("HK_HD", "__global__", "__device__", "template")):
                m = re.search(r"(\w+)\(", l)
                if m:
                    out.append((n, m.group(1)))
        return out

    csrc = os.path.join(ROOT, "hockey_env_b200", "csrc")
    fmap = {f: funcs(os.path.join(csrc, f)) for f in os.listdir(csrc)}
    with open(os.path.join(PROF, TAG + "_hotspots_by_function.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "file", "function", "pct_warp_instructions", "pct_stall_samples", "active_threads_per_instruction"])
        for kern, agg in data.items():
            tot = sum(a[0] for a in agg.values())
            tots = sum(a[2] for a in agg.values())
            byfn = {}
            for (fl, l), a in agg.items():
                name = "?"
                if fl in fmap and fmap[fl]:
                    starts = [x[0] for x in fmap[fl]]
                    k = bisect.bisect_right(starts, l) - 1
                    if k >= 0:
                        name = fmap[fl][k][1]
                b = byfn.setdefault((fl, name), [0, 0, 0])
                b[0] += a[0]
                b[1] += a[1]
                b[2] += a[2]
            for (fl, n), b in sorted(byfn.items(), key=lambda kv: -kv[1][2])[:30]:
                w.writerow([kern, fl, n, "%.2f" % (100 * b[0] / tot), "%.2f" % (100 * b[2] / max(tots, 1)), "%.1f" % (b[1] / max(b[0], 1))])
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(d["Kernel Name"][:40], "ms", d["gpu__time_duration.sum"], "dram MB", d["dram__bytes_read.sum"], "+", d["dram__bytes_write.sum"],
              "inst", d["smsp__inst_executed.sum"], "lanes/inst", d["smsp__thread_inst_executed_per_inst_executed.ratio"],
              "issue%", d["smsp__issue_active.avg.pct_of_peak_sustained_active"],
              "no_inst", d["smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"],
              "barrier", d["smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"], "regs", d["launch__registers_per_thread"])


if __name__ == "__main__":
    main()

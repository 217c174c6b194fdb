"""Round-2 evidence: turns gpurun_out/<tag>_full.ncu-rep + the bench lines of scripts/evidence_r2.sh into the tracked files
under profiles/ and (re)writes profiles/roofline_capture.json, which bench.py reads for `roofline.traffic` and
`roofline.issue` (ncu numbers of the steady-state tick of THIS build; the times next to them are measured live).
Run in the build container (ncu reads reports without a GPU):  python scripts/extract_profiles_r2.py <tag> [config]"""
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2e"
CONFIG = sys.argv[2] if len(sys.argv) > 2 else "normal65k"
sys.argv = [sys.argv[0], TAG]
csv.field_size_limit(10 ** 9)


def main():
    os.makedirs(PROF, exist_ok=True)
    for f in glob.glob(os.path.join(OUT, TAG + "_bench_*.json")) + glob.glob(os.path.join(OUT, TAG + "_launches*.csv")) + \
            glob.glob(os.path.join(OUT, TAG + "_phase_cycles.txt")) + glob.glob(os.path.join(OUT, TAG + "_pytest.txt")):
        shutil.copy(f, os.path.join(PROF, os.path.basename(f)))
    rep = os.path.join(OUT, TAG + "_full.ncu-rep")
    if not os.path.exists(rep):
        print("no", rep)
        return
    import extract_profiles as E  # raw metrics + per-function hot spots (same extraction as round 1)
    E.TAG, E.REP, E.SHORT = TAG, rep, TAG
    E.main()
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    ks = [dict(zip(hdr, r)) for r in rows[2:]]
    unit = dict(zip(hdr, units))

    def mb(d, k):
        v = float(d[k])
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[unit[k]]
    dom = max(ks, key=lambda d: float(d["gpu__time_duration.sum"]))
    tus = {"us": 1e-3, "ns": 1e-6, "ms": 1.0}[unit["gpu__time_duration.sum"]]
    cap_path = os.path.join(PROF, "roofline_capture.json")
    cap = json.load(open(cap_path)) if os.path.exists(cap_path) else {}
    cap[CONFIG] = {
        "source": f"profiles/{TAG}_raw_metrics.csv (ncu --set full --clock-control none, one steady-state tick after a 400-tick pre-roll)",
        "kernel": dom["Kernel Name"].split("(")[0].replace("void <unnamed>::", ""),
        "dram_bytes_per_launch": mb(dom, "dram__bytes_read.sum") + mb(dom, "dram__bytes_write.sum"),
        "warp_inst_per_tick": sum(float(d["smsp__inst_executed.sum"]) for d in ks),
        "capture_ms_per_tick": sum(float(d["gpu__time_duration.sum"]) * tus for d in ks),
        "kernels": {d["Kernel Name"].split("(")[0].replace("void <unnamed>::", ""): {
            "ms": float(d["gpu__time_duration.sum"]) * tus, "warp_inst": float(d["smsp__inst_executed.sum"]),
            "dram_bytes": mb(d, "dram__bytes_read.sum") + mb(d, "dram__bytes_write.sum"),
            "lanes_per_inst": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]),
            "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
            "registers": int(float(d["launch__registers_per_thread"]))} for d in ks},
    }
    json.dump(cap, open(cap_path, "w"), indent=1)
    print(json.dumps(cap[CONFIG], indent=1))


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    main()

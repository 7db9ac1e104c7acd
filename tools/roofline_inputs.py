"""Extracts what bench.py's roofline objects quote from ncu -- DRAM traffic, executed warp instructions and issue-slot
utilisation of the dominant kernels -- out of the .ncu-rep captures of tools/profile_r2.sh, into
profiles/r2_roofline_inputs.json (tracked; bench.py reads it).  Usage: roofline_inputs.py"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
        "warp_instructions": "smsp__inst_executed.sum",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "duration": "gpu__time_duration.sum"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "msecond": 1e3,
        "nsecond": 1e-3}


def kernels(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")]}
        for key, metric in WANT.items():
            v = float(r[h.index(metric)].replace(",", ""))
            d[key] = v * UNIT.get(units[h.index(metric)], 1)
        res.append(d)
    return res


def main():
    out = {"source": "ncu --set full --clock-control none captures of tools/profile_r2.sh (cold cache, serialised)"}
    for name, rep in (("sweep", "prof_sweep_r2"), ("dwa_c4", "prof_c4_r2"), ("mirror", "prof_mirror_r2"),
                      ("obstacle", "prof_obstacle_r2")):
        path = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        out[name] = [{"kernel": k["kernel"].split("(")[0].replace("void ", "").replace("navgpu::", ""),
                      "dram_bytes": int(k["dram_read"] + k["dram_write"]),
                      "warp_instructions": int(k["warp_instructions"]), "issue_active_pct": round(k["issue_active_pct"], 2),
                      "ncu_duration_us": round(k["duration"], 2)} for k in kernels(path)]
    with open(os.path.join(ROOT, "profiles", "r2_roofline_inputs.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

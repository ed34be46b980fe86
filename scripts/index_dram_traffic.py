#!/usr/bin/env python
"""Build profiles/dram_traffic.json — the index bench.py reads `roofline.traffic` from.

Each entry names one `ncu --set full` capture of the fused sweep kernel (a summary written by
scripts/ncu_summary.py) and the configuration it was taken on; the value is
dram__bytes_read.sum + dram__bytes_write.sum of ONE launch. Edit CAPTURES when a new capture is
committed, then run this script.
"""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
# (workload, rows, cols, fused iterations, summary file under profiles/)
CAPTURES = [
    # round 2, final kernels (scripts/gpu/ncu_captures.sh)
    ("jacobi5", 16384, 16384, 6, "r02_final_ncu_jacobi5_summary.txt"),
    ("hotspot", 16384, 16384, 4, "r02_final_ncu_hotspot_summary.txt"),
    ("fdtd", 4608, 4608, 4, "r02_final_ncu_fdtd_summary.txt"),
    ("convection_pt", 4096, 8192, 1, "r02_final_ncu_convection_pt_summary.txt"),
    ("jacobi_r3", 16384, 16384, 2, "r02_final_ncu_jacobi_r3_summary.txt"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def first_launch_metrics(path: Path) -> dict:
    values = {}
    for line in path.read_text().splitlines():
        parts = line.split()
        if line.startswith("kernel:") and values:
            break
        if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            values[parts[0]] = float(parts[1]) * SCALE.get(parts[2], 1.0)
        if len(parts) == 3 and parts[0] == "gpu__time_duration.sum":
            values["time_us"] = float(parts[1]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[parts[2]]
    return values


def main():
    entries = []
    for workload, rows, cols, k, name in CAPTURES:
        path = ROOT / "profiles" / name
        if not path.exists():
            continue
        m = first_launch_metrics(path)
        entries.append({
            "workload": workload, "rows": rows, "cols": cols, "fused_iterations": k,
            "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
            "dram_bytes_read": m["dram__bytes_read.sum"], "dram_bytes_write": m["dram__bytes_write.sum"],
            "launch_time_us_under_ncu": m.get("time_us"), "source": f"profiles/{name}",
        })
    out = ROOT / "profiles" / "dram_traffic.json"
    out.write_text(json.dumps(entries, indent=1) + "\n")
    print(f"{out}: {len(entries)} captures")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Summarise an `ncu --set full --import-source on` capture of the walk kernels into the tracked files the bench line and DESIGN.md cite.

    python profiles/ncu_extract.py gpurun_out/<capture>.ncu-rep <commit> [round-tag]

Writes profiles/<tag>_ncu_walk_kernels.json (per-kernel metrics), profiles/<tag>_walk_top_stalls.txt (hottest source lines by
executed warp instructions and by stall samples) and profiles/ncu_traffic.json (DRAM bytes of the walk kernels per walk launch,
with the commit the capture was taken at — bench.py reports it as `roofline.traffic`)."""

from __future__ import annotations

import csv
import datetime
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def ncu(*args) -> str:
    return subprocess.run(["ncu", *args], capture_output=True, text=True, check=True).stdout


def to_bytes(value: str, unit: str) -> float:
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value) * scale.get(unit, 1.0)


def main() -> None:
    rep, commit = sys.argv[1], sys.argv[2]
    tag = sys.argv[3] if len(sys.argv) > 3 else "r02"
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    head, units, data = rows[0], rows[1], rows[2:]
    kernels = []
    for r in data:
        k = {"Kernel Name": r[head.index("Kernel Name")]}
        for w in WANT:
            if w in head:
                i = head.index(w)
                k[w] = f"{r[i]} {units[i]}".strip()
        k["_dram_bytes"] = to_bytes(r[head.index("dram__bytes_read.sum")], units[head.index("dram__bytes_read.sum")]) + \
            to_bytes(r[head.index("dram__bytes_write.sum")], units[head.index("dram__bytes_write.sum")])
        k["_ms"] = float(r[head.index("gpu__time_duration.sum")]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(units[head.index("gpu__time_duration.sum")], 1.0)
        kernels.append(k)
    big = [k for k in kernels if k["_ms"] > 0.05]
    (ROOT / "profiles" / f"{tag}_ncu_walk_kernels.json").write_text(json.dumps({"commit": commit, "capture": Path(rep).name, "kernels": kernels}, indent=1) + "\n")
    traffic = sum(k["_dram_bytes"] for k in big)
    (ROOT / "profiles" / "ncu_traffic.json").write_text(json.dumps({
        "walk_dram_bytes_per_launch": traffic, "commit": commit, "captured": datetime.date.today().isoformat(),
        "kernels": {k["Kernel Name"][:60]: {"ms": round(k["_ms"], 3), "dram_bytes": k["_dram_bytes"]} for k in big},
        "how": "ncu --set full --clock-control none --import-source on -k regex:walk_ (bench.py --workload L); dram__bytes_read.sum + dram__bytes_write.sum of the walk kernels of one walk launch"},
        indent=1) + "\n")
    # hottest source lines of every long kernel
    out = [f"Hottest CUDA source lines of the walk kernels — {Path(rep).name} at commit {commit} (ncu --page source --print-source cuda,sass; share of executed warp instructions / of stall samples)"]
    seen = set()
    for k in big:
        name = k["Kernel Name"]
        short = name.split("(")[0].split("<")[0].replace("void ", "").strip()
        if short in seen:
            continue
        seen.add(short)
        src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{short}", "--launch-count", "1"))))
        fn, agg = None, []
        for r in src:
            if len(r) >= 2 and r[0] == "File Path":
                fn = r[1].split("/")[-1]
            if len(r) > 8 and r[0].isdigit():
                try:
                    agg.append((fn, int(r[0]), r[1].strip(), int(r[7] or 0), int(r[4] or 0)))
                except ValueError:
                    pass
        ti, ts = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
        out.append(f"\n== {name[:100]}  ({k['_ms']:.2f} ms, {k.get('smsp__inst_executed.sum', '')})")
        for a in sorted(agg, key=lambda a: -a[3])[:14]:
            out.append(f"  {100 * a[3] / ti:5.1f}% instr  {100 * a[4] / ts:5.1f}% stalls  {a[0]}:{a[1]}  {a[2][:110]}")
    (ROOT / "profiles" / f"{tag}_walk_top_stalls.txt").write_text("\n".join(out) + "\n")
    print(f"walk DRAM bytes per launch: {traffic / 1e9:.2f} GB over {len(big)} kernels")


if __name__ == "__main__":
    main()

"""Device timing of abb_group_union_host (the effective-reach reduction) on synthetic arrays.

    python profiles/union_microbench.py [n_groups] > profiles/r01_union_microbench.json

CUDA-event time of the kernels (count → scan → fill → radix sort → unique → histogram → scan), as the library
reports it; host↔device copies are outside that window.  Shape: every group (vulnerability) has 1–3 members (servers),
every member 8–40 items (tool / credential / agent label ranks) drawn from a small pool so unions really de-duplicate."""

from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from agent_bom_b200.engine import group_union  # noqa: E402


def main() -> None:
    n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    n_members = n_groups // 2
    rng = np.random.default_rng(5)
    icount = rng.integers(8, 41, n_members)
    ioff = np.concatenate([[0], np.cumsum(icount)]).astype(np.int64)
    items = rng.integers(0, 4096, int(ioff[-1])).astype(np.int32) | (rng.integers(0, 3, int(ioff[-1])).astype(np.int32) << 29)
    mcount = rng.integers(1, 4, n_groups)
    moff = np.concatenate([[0], np.cumsum(mcount)]).astype(np.int64)
    members = rng.integers(0, n_members, int(moff[-1])).astype(np.int32)
    w0 = rng.integers(0, 8, n_members).astype(np.uint8)
    w1 = rng.integers(0, 4, n_members).astype(np.uint8)
    pairs = int((ioff[members + 1] - ioff[members]).sum())
    times, walls = [], []
    for _ in range(6):
        t = time.perf_counter()
        off, out, g0, g1, ms = group_union(moff, members, ioff, items, w0, w1)
        walls.append((time.perf_counter() - t) * 1e3)
        times.append(ms)
    ms = float(np.median(times[2:]))
    print(json.dumps({
        "kernel": "abb_group_union_host (union_count + union_fill + cub radix sort / unique / scan + histogram)",
        "groups": n_groups, "members": n_members, "pairs": pairs, "unique_items_out": int(off[-1]),
        "device_ms": ms, "pairs_per_s": pairs / (ms * 1e-3), "groups_per_s": n_groups / (ms * 1e-3),
        "algorithmic_bytes": {"per_pair": 24, "note": "4 B item read + 12 B key/value written by the expansion, 8 B key read by the histogram; the radix passes in between are the sort's own traffic"},
        "achieved_GBps": pairs * 24 / (ms * 1e-3) / 1e9, "end_to_end_ms_with_copies": float(np.median(walls[2:])),
    }))


if __name__ == "__main__":
    main()

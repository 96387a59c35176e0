#!/usr/bin/env python3
"""Chunked host walks at L scale: parity against the one-piece call for every finding, and the end-to-end time of
`DeviceGraph.exposure_many` (the call bench.py's `e2e` times) for 1 / 2 / 4 / 8 ranges on the same box.

    python profiles/chunk_check.py [--workload L] > gpurun_out/r02_chunk_check.json"""

from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="L")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sample", type=int, default=150_000)
    args = ap.parse_args()
    from agent_bom_b200 import estate
    from agent_bom_b200.engine import DeviceGraph
    from agent_bom_b200.graph import csr as csrmod
    from bench import WORKLOADS

    est = estate.generate(WORKLOADS[args.workload][0], 2145, estate.BENCH_KNOBS, exact_rank=False)
    host = csrmod.from_arrays(None, est.node_type, est.src, est.dst, est.rel, est.flags, node_rank=est.node_rank)
    dg = DeviceGraph.upload(host)
    findings = np.ascontiguousarray(est.findings, dtype=np.int32)
    nq = len(findings)
    dg.set_option("chunk_min", min(2 << 20, max(64, nq // 2)))
    rng = np.random.default_rng(3)
    sample = np.sort(rng.choice(nq, size=min(args.sample, nq), replace=False))
    out = {"workload": args.workload, "findings": nq, "runs": []}
    ref = None
    ok_all = True
    for chunks in (1, 2, 4, 8):
        dg.set_option("chunks", chunks)
        for _ in range(2):
            w, p = dg.exposure_many(findings, 4, zero_copy=True)
        used = dg.get_option("last_host_chunks")
        times = []
        for _ in range(args.reps):
            del w, p
            t0 = time.perf_counter()
            w, p = dg.exposure_many(findings, 4, zero_copy=True)
            times.append(time.perf_counter() - t0)
        run = {"chunks": chunks, "ranges_used": int(used), "e2e_ms_best": 1e3 * min(times), "e2e_ms_median": 1e3 * float(np.median(times)),
               "traversals_per_s": nq / float(np.median(times)), "result_nodes": int(w.nodes.shape[0]), "d2h_bytes": int(w.d2h_bytes + p.d2h_bytes)}
        if ref is None:
            ref = (w.count.copy(), w.maxd.copy(), np.array(w.hist), (w.flags & 2).copy(), [w.slice(int(q)).copy() for q in sample], p.off.copy())
        else:
            same = (np.array_equal(w.count, ref[0]) and np.array_equal(w.maxd, ref[1]) and np.array_equal(w.hist, ref[2]) and np.array_equal(w.flags & 2, ref[3])
                    and np.array_equal(p.off, ref[5]) and all(np.array_equal(w.slice(int(q)), s) for q, s in zip(sample, ref[4])))
            run["equals_one_piece"] = bool(same)
            run["checked"] = f"count / max depth / histogram / flags of all {nq:,} findings, node slices of {len(sample):,} sampled findings, path-row offsets"
            ok_all = ok_all and same
        out["runs"].append(run)
        print(f"[chunk_check] {run}", file=sys.stderr, flush=True)
    out["all_equal"] = ok_all
    print(json.dumps(out), flush=True)
    return 0 if ok_all else 1


if __name__ == "__main__":
    raise SystemExit(main())

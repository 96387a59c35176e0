"""CPU baseline of the UNMODIFIED Python reference for the hot path (BASELINE.md §3 engines 1 and 2).

Runs only where /root/reference exists (the build container): the reference cannot travel to the GPU box, so its
numbers are measured here, written to ``profiles/r02_reference_python.json`` and carried into bench.py's JSON line
(``cpu_baseline.reference_python``).  Timing idiom: ``time.perf_counter`` around the plain calls, as
``/root/reference/scripts/run_scale_evidence.py:37-61`` does.

    python oracle/time_reference_python.py [--agents 2000 30000] [--paths-agents 60] [--procs 8]

One *traversal* = ``UnifiedGraph.impact_of(f, max_depth=4)`` + f's share of ``_derived_attack_paths(graph)``
(reference graph/container.py:230-279, api/routes/graph.py:686-786).  ``_derived_attack_paths`` rebuilds an O(|E|) pair
map per emitted path (routes/graph.py:492-497), so it is timed on a small estate and reported per finding with the edge
count next to it — on the benchmark estates it would run for hours.

Test / measurement infrastructure — nothing in the product imports this."""

from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(ROOT))

from agent_bom.api.routes.graph import _derived_attack_paths  # noqa: E402
from agent_bom.graph import build_unified_graph_from_report  # noqa: E402
from agent_bom.graph.types import EntityType  # noqa: E402

from agent_bom_b200 import estate  # noqa: E402

_G = None
_F: list = []


def _shard(args):
    lo, hi = args
    g, f = _G, _F
    t = time.perf_counter()
    n = 0
    for i in range(lo, hi):
        n += g.impact_of(f[i])["affected_count"]
    return time.perf_counter() - t, n


def findings_of(g):
    return [n.id for n in g.nodes.values() if n.entity_type in (EntityType.VULNERABILITY, EntityType.MISCONFIGURATION)]


def time_impact(agents: int, procs: int, budget_s: float) -> dict:
    global _G, _F
    est = estate.generate(agents, 2145, estate.BENCH_KNOBS, exact_rank=False)
    t = time.perf_counter()
    g = build_unified_graph_from_report(est.report_json())
    t_build = time.perf_counter() - t
    f = findings_of(g)
    # single process: a bounded, evenly spaced sample
    probe = f[:: max(1, len(f) // 500)][:500]
    t = time.perf_counter()
    for x in probe:
        g.impact_of(x)
    rate = len(probe) / (time.perf_counter() - t)
    k = int(min(len(f), max(len(probe), rate * budget_s)))
    sample = f[:: max(1, len(f) // k)][:k]
    t = time.perf_counter()
    reached = 0
    for x in sample:
        reached += g.impact_of(x)["affected_count"]
    dt = time.perf_counter() - t
    out = {
        "agents": agents, "nodes": len(g.nodes), "edges": len(g.edges), "findings": len(f), "reference_builder_s": round(t_build, 2),
        "impact_of": {"sample": len(sample), "seconds": round(dt, 3), "findings_per_s_single_process": round(len(sample) / dt, 1),
                      "mean_reach": round(reached / max(1, len(sample)), 1)},
    }
    # fork pool over source shards (graph shared copy-on-write) — BASELINE.md engine 2
    _G, _F = g, sample
    shards = [(len(sample) * i // procs, len(sample) * (i + 1) // procs) for i in range(procs)]
    ctx = mp.get_context("fork")
    t = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_shard, shards)
    dtp = time.perf_counter() - t
    out["impact_of"]["fork_pool"] = {"processes": procs, "seconds": round(dtp, 3), "findings_per_s": round(len(sample) / dtp, 1),
                                    "slowest_shard_s": round(max(r[0] for r in res), 3)}
    return out


def time_paths(agents: int) -> dict:
    est = estate.generate(agents, 2145, estate.BENCH_KNOBS, exact_rank=True)
    g = build_unified_graph_from_report(est.report_json())
    f = findings_of(g)
    t = time.perf_counter()
    paths = _derived_attack_paths(g)
    dt = time.perf_counter() - t
    t = time.perf_counter()
    for x in f:
        g.impact_of(x)
    di = time.perf_counter() - t
    return {"agents": agents, "nodes": len(g.nodes), "edges": len(g.edges), "findings": len(f), "paths": len(paths), "derived_attack_paths_s": round(dt, 3),
            "findings_per_s": round(len(f) / dt, 2), "impact_of_s": round(di, 3),
            "traversals_per_s_single_process": round(len(f) / (dt + di), 2)}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--agents", type=int, nargs="*", default=[2000, 30000])
    ap.add_argument("--paths-agents", type=int, nargs="*", default=[30, 60])
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--budget", type=float, default=20.0, help="seconds of single-process impact_of per estate")
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r02_reference_python.json"))
    args = ap.parse_args()
    doc = {"what": "unmodified Python reference (/root/reference/src, v0.87.1) timed in the build container",
           "host": {"cpu_count": os.cpu_count(), "python": sys.version.split()[0]},
           "estate_knobs": "agent_bom_b200.estate.BENCH_KNOBS, seed 2145", "impact": [], "derived_paths": []}
    for a in args.paths_agents:
        doc["derived_paths"].append(time_paths(a))
        print(json.dumps(doc["derived_paths"][-1]), flush=True)
    for a in args.agents:
        doc["impact"].append(time_impact(a, args.procs, args.budget))
        print(json.dumps(doc["impact"][-1]), flush=True)
    Path(args.out).write_text(json.dumps(doc, indent=1) + "\n")


if __name__ == "__main__":
    main()

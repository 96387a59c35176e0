"""Measurement for SURVEY §8 f1, report-JSON half (runs only where /root/reference exists): the reference's
``build_unified_graph_from_report`` against this repo's record builder and its columnar builder on the same report.

    python oracle/time_builder.py [agents]        # default 600 agents, dense bench knobs → profiles/r02_builder_timing.json

Test / measurement infrastructure — nothing in the product imports this."""

from __future__ import annotations

import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(ROOT))

from agent_bom.graph import build_unified_graph_from_report as ref_build  # noqa: E402

from agent_bom_b200 import estate  # noqa: E402
from agent_bom_b200.graph import build_unified_graph_from_report, csr as csrmod  # noqa: E402


def best(fn, n=3):
    out, t_best = None, 1e9
    for _ in range(n):
        t = time.perf_counter()
        out = fn()
        t_best = min(t_best, time.perf_counter() - t)
    return out, t_best


def main() -> None:
    agents = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    est = estate.generate(agents, 2145, estate.BENCH_KNOBS, exact_rank=True)
    rep = est.report_json()
    g_ref, t_ref = best(lambda: ref_build(rep), 1)
    ref_csr, t_ref_csr = best(lambda: csrmod.from_unified_graph(g_ref), 1)
    (g_rec, rec_csr), t_rec = best(lambda: (lambda g: (g, g.csr))(build_unified_graph_from_report(rep)))
    (g_col, col_csr), t_col = best(lambda: (lambda g: (g, g.csr))(build_unified_graph_from_report(rep, columnar=True)))
    same = all((getattr(ref_csr, k) == getattr(col_csr, k)).all() for k in ("fwd_off", "fwd_nbr", "fwd_meta", "rev_off", "rev_nbr", "rev_meta", "node_type", "node_rank"))
    doc = {"agents": agents, "nodes": len(g_ref.nodes), "edges": len(g_ref.edges), "reference_builder_s": round(t_ref, 3), "reference_records_to_csr_s": round(t_ref_csr, 3),
           "record_builder_plus_csr_s": round(t_rec, 3), "columnar_builder_plus_csr_s": round(t_col, 3),
           "ms_per_agent": {"reference": round(1000 * (t_ref + t_ref_csr) / agents, 2), "records": round(1000 * t_rec / agents, 2), "columnar": round(1000 * t_col / agents, 2)},
           "speedup_columnar_vs_reference": round((t_ref + t_ref_csr) / t_col, 1), "records_synthesised_by_columnar_build": len(g_col.nodes._cache),
           "csr_identical_to_reference_graph": bool(same), "host": "build container (8 cores)"}
    print(json.dumps(doc))
    (ROOT / "profiles" / "r02_builder_timing.json").write_text(json.dumps(doc, indent=1) + "\n")


if __name__ == "__main__":
    main()

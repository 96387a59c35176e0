"""Measurement for SURVEY §8 f1 (runs only where /root/reference exists): the reference's load_graph against
agent_bom_b200.graph.snapshot.load_snapshot on the same SQLite file, written by the reference's save_graph.

    python oracle/time_snapshot_load.py [agents]      # default 600 agents, dense bench knobs

Test / measurement infrastructure — nothing in the product imports this."""

from __future__ import annotations

import sqlite3
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(ROOT))

from agent_bom.db import graph_store as ref_db  # noqa: E402
from agent_bom.graph import build_unified_graph_from_report  # noqa: E402

from agent_bom_b200 import estate  # noqa: E402
from agent_bom_b200.graph import csr as csrmod  # noqa: E402
from agent_bom_b200.graph.snapshot import load_snapshot  # noqa: E402


def main() -> None:
    agents = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    est = estate.generate(agents, 2145, estate.BENCH_KNOBS, exact_rank=True)
    t = time.perf_counter()
    g = build_unified_graph_from_report(est.report_json())
    t_build = time.perf_counter() - t
    g.scan_id, g.tenant_id = "scan-1", "default"
    path = Path(tempfile.mkdtemp()) / "graph.sqlite"
    conn = sqlite3.connect(str(path))
    conn.row_factory = sqlite3.Row
    ref_db._init_db(conn)
    ref_db.save_graph(conn, g)
    conn.commit()
    t = time.perf_counter()
    loaded = ref_db.load_graph(conn, tenant_id="default", scan_id="scan-1")
    t_ref_load = time.perf_counter() - t
    t = time.perf_counter()
    ref_csr = csrmod.from_unified_graph(loaded)
    t_ref_csr = time.perf_counter() - t
    conn.close()
    t = time.perf_counter()
    snap = load_snapshot(path, tenant_id="default", scan_id="scan-1")
    t_snap = time.perf_counter() - t
    same = all((getattr(ref_csr, k) == getattr(snap.csr, k)).all() for k in ("fwd_off", "fwd_nbr", "fwd_meta", "rev_off", "rev_nbr", "rev_meta", "node_type", "node_rank"))
    print(f"{agents} agents: {len(loaded.nodes)} nodes / {len(loaded.edges)} edges; reference builder {t_build:.2f}s; "
          f"reference load_graph {t_ref_load:.2f}s (+{t_ref_csr:.2f}s records->CSR); load_snapshot (CSR included) {t_snap:.2f}s; "
          f"speed-up {(t_ref_load + t_ref_csr) / t_snap:.1f}x; CSR identical: {same}")


if __name__ == "__main__":
    main()

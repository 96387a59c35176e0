"""Topology-first snapshot loader against what the reference's own load_graph returns.

``tests/golden/snapshot/graph.sqlite.gz`` was written by the reference's ``save_graph`` (``oracle/make_golden.py
--snapshot-only``): two tenants, a re-saved scan (INSERT OR REPLACE moves rows), dangling edges, a blank-tenant save and
materialised attack paths.  ``expected.json.gz`` is the reference's ``load_graph`` for each (tenant, scan) request.
"""

from __future__ import annotations

import gzip
import json
from pathlib import Path

import numpy as np
import pytest

from agent_bom_b200.graph import csr as csrmod
from agent_bom_b200.graph.schema import ENTITY_CODE, REL_CODE, enum_value
from agent_bom_b200.graph.snapshot import SnapshotGraph, load_snapshot, normalize_graph_tenant_id

GOLD = Path(__file__).parent / "golden" / "snapshot"


@pytest.fixture(scope="module")
def db_path(tmp_path_factory) -> Path:
    path = tmp_path_factory.mktemp("snap") / "graph.sqlite"
    path.write_bytes(gzip.decompress((GOLD / "graph.sqlite.gz").read_bytes()))
    return path


CASES = json.loads(gzip.decompress((GOLD / "expected.json.gz").read_bytes()))


@pytest.mark.parametrize("case", CASES, ids=[f"{c['tenant'] or '-'}:{c['scan'] or 'latest'}" for c in CASES])
def test_snapshot_matches_reference_load_graph(db_path, case):
    g = load_snapshot(db_path, tenant_id=case["tenant"], scan_id=case["scan"])
    assert isinstance(g, SnapshotGraph)
    assert g.scan_id == case["scan_id"] and g.tenant_id == normalize_graph_tenant_id(case["tenant"]) == case["tenant_id"]
    if case["nodes"]:
        assert g.created_at == case["created_at"]
    # node order, kinds and the resident columns
    got_nodes = [[n.id, enum_value(n.entity_type), n.label, n.severity, float(n.risk_score or 0.0)] for n in g.nodes.values()]
    assert got_nodes == case["nodes"]
    # the edge stream in the reference's row order (adjacency order fixes BFS discovery order)
    got_edges = [[e.source, e.target, enum_value(e.relationship), e.direction, bool(e.traversable), float(e.weight)] for e in g.edges]
    assert got_edges == case["edges"]
    assert len(g.edges) == len(case["edges"])
    # the CSR equals the one built from the reference-shaped records
    c = g.csr
    assert c.node_ids == [n[0] for n in case["nodes"]] and c.n_real == len(case["nodes"])
    assert c.node_type.tolist() == [ENTITY_CODE[n[1]] for n in case["nodes"]]
    idx = {n[0]: i for i, n in enumerate(case["nodes"])}
    if case["edges"]:
        want = csrmod.from_arrays(
            c.node_ids, c.node_type, [idx[e[0]] for e in case["edges"]], [idx[e[1]] for e in case["edges"]], [REL_CODE[e[2]] for e in case["edges"]],
            [(1 if e[4] else 0) | (2 if e[3] == "bidirectional" else 0) for e in case["edges"]], n_real=c.n_real)
        for name in ("fwd_off", "fwd_nbr", "fwd_meta", "fwd_eid", "rev_off", "rev_nbr", "rev_meta", "rev_eid", "node_rank"):
            assert np.array_equal(getattr(c, name), getattr(want, name)), name
    assert [p.to_dict() for p in g.attack_paths] == case["attack_paths"]
    assert [r.to_dict() for r in g.interaction_risks] == case["interaction_risks"]


@pytest.mark.parametrize("case", [c for c in CASES if c["nodes"]], ids=[f"{c['tenant'] or '-'}:{c['scan'] or 'latest'}" for c in CASES if c["nodes"]])
def test_lazy_records_hydrate_to_the_reference_dicts(db_path, case):
    g = load_snapshot(db_path, tenant_id=case["tenant"], scan_id=case["scan"])
    for want in case["node_dicts"]:
        node = g.nodes[want["id"]]
        assert node._full is None                       # nothing wide was read at load time
        got = node.to_dict()
        assert json.loads(json.dumps(got, sort_keys=True, default=str)) == want
        assert node.attributes == want["attributes"] and node.compliance_tags == want["compliance_tags"]
    by_key = {(e[0], e[1], e[2]): i for i, e in enumerate(case["edges"])}
    for want in case["edge_dicts"]:
        edge = g.edges[by_key[(want["source"], want["target"], want["relationship"])]]
        assert edge.id == want["id"]
        assert json.loads(json.dumps(edge.to_dict(), sort_keys=True, default=str)) == want
        if edge.is_bidirectional:
            twin = g.edges[by_key[(want["source"], want["target"], want["relationship"])]].reversed_copy()
            assert (twin.source, twin.target) == (want["target"], want["source"]) and twin.evidence == want["evidence"]


def test_snapshot_graph_is_immutable_and_accepts_a_connection(db_path):
    import sqlite3

    conn = sqlite3.connect(str(db_path))
    try:
        g = load_snapshot(conn, tenant_id="default", scan_id="scan-b")
    finally:
        conn.close()
    assert len(g.nodes) == 1342 and len(g.edges) == 4135
    with pytest.raises(TypeError):
        g.add_edge(g.edges[0])
    with pytest.raises(TypeError):
        g.add_node(next(iter(g.nodes.values())))
    assert g.edges[-1].source == g.edges[len(g.edges) - 1].source
    assert [e.id for e in g.edges[2:5]] == [g.edges[i].id for i in (2, 3, 4)]
    # host CSR point queries work off the thin records
    first = next(iter(g.nodes))
    assert [e.source for e in g.edges_from(first)] == [first] * len(g.edges_from(first))


def test_missing_snapshot_is_an_empty_graph(db_path):
    g = load_snapshot(db_path, tenant_id="nobody")
    assert g.scan_id == "" and not g.nodes and len(g.edges) == 0 and g.csr.n_nodes == 0


class _SqliteInner:
    """Stands in for the reference's SQLiteGraphStore (api/graph_store.py:259-263): the attribute the store looks for."""

    def __init__(self, path):
        self._db_path = path
        self.load_calls = 0

    def load_graph(self, **_kw):
        self.load_calls += 1
        return None


def test_store_loads_cache_misses_topology_first(db_path):
    """No GPU involved: materialised attack paths page straight out of the snapshot, in the stores' order."""
    from agent_bom_b200.store import B200GraphStore

    inner = _SqliteInner(db_path)
    store = B200GraphStore(inner=inner)
    sid, created, page, total = store.attack_paths(tenant_id="default", scan_id="scan-b", offset=0, limit=5)
    case = next(c for c in CASES if c["tenant"] == "default" and c["scan"] == "scan-b")
    want = sorted(case["attack_paths"], key=lambda p: (-p["composite_risk"], p["source"], p["target"]))
    assert (sid, created, total) == ("scan-b", case["created_at"], len(want)) and [p.to_dict() for p in page] == want[:5]
    assert isinstance(store.load_graph(tenant_id="default", scan_id="scan-b"), SnapshotGraph) and inner.load_calls == 0
    assert store.impact_of(tenant_id="default", scan_id="scan-b", node_id="no-such-node") is None
    assert store.attack_paths(tenant_id="nobody") == ("", "", [], 0) and inner.load_calls == 1      # falls through to the inner store, which has nothing

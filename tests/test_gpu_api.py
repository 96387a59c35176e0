"""GPU: the Python drop-in surface (UnifiedGraph / store / backend / dependency reach / derived paths) against the
unmodified reference's answers in the golden fixtures — string ids, dict shapes and list orders as the reference returns them."""

from __future__ import annotations

import pytest

from golden_util import ALL_FIXTURES, SMALL_FIXTURES, graph_from_fixture, load

pytestmark = pytest.mark.gpu

_graphs: dict = {}


def graph(name):
    if name not in _graphs:
        _graphs[name] = graph_from_fixture(load(name))
    return _graphs[name]


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_impact_of_dicts(name):
    doc, g = load(name), graph(name)
    ids = doc["node_ids"]
    cases = doc["cases"]["impact"]
    for depth in sorted({c["d"] for c in cases}):
        sub = [c for c in cases if c["d"] == depth]
        got = g.impact_of_many([ids[c["s"]] for c in sub], depth)
        for c, r in zip(sub, got):
            assert r == {"node_id": ids[c["s"]], "affected_nodes": sorted(ids[i] for i in c["nodes"]), "affected_by_type": c["by_type"],
                         "affected_count": c["count"], "max_depth_reached": c["maxd"]}
    assert g.impact_of("no-such-node") == doc["cases"]["impact_missing"]


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_bfs_paths(name):
    doc, g = load(name), graph(name)
    ids = doc["node_ids"]
    cases = doc["cases"]["bfs"]
    for key in sorted({(c["d"], c["t"]) for c in cases}):
        sub = [c for c in cases if (c["d"], c["t"]) == key]
        got = g.bfs_many([ids[c["s"]] for c in sub], key[0], key[1])
        for c, paths in zip(sub, got):
            assert paths == [[ids[i] for i in p] for p in c["paths"]]
    assert g.bfs("no-such-node") == []


@pytest.mark.parametrize("name", SMALL_FIXTURES + ["mesh_inventory"])
def test_reachable_and_shortest_path(name):
    doc, g = load(name), graph(name)
    ids = doc["node_ids"]
    for c in doc["cases"]["reachable"]:
        got = g.reachable_from(ids[c["s"]], c["d"], traversable_only=c["t"], include_source=c["inc"])
        assert got == {ids[i] for i in c["nodes"]}
    for c in doc["cases"]["shortest"][:400]:
        got = g.shortest_path(ids[c["a"]], ids[c["b"]])
        assert got == (None if c["path"] is None else [ids[i] for i in c["path"]])
    assert g.shortest_path("nope", ids[0]) is None and g.reachable_from("nope") == set()


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_traverse_subgraph(name):
    from agent_bom_b200.graph import RelationshipType
    from agent_bom_b200.graph.schema import REL_CODE, RELATIONSHIP_VALUES, enum_value

    doc, g = load(name), graph(name)
    ids = doc["node_ids"]
    cases = doc["cases"]["traverse"]
    for c in cases[:: max(1, len(cases) // 400)]:
        kw = dict(c["kw"])
        if "relationship_types" in kw:
            kw["relationship_types"] = {RelationshipType(RELATIONSHIP_VALUES[r]) for r in kw["relationship_types"]}
        roots = [ids[r] if r >= 0 else "missing:root" for r in c["roots"]]
        call = dict(direction=c["direction"], max_depth=4, max_nodes=500, max_edges=10_000)
        call.update(kw)
        sub, depth_by, truncated = g.traverse_subgraph(roots, **call)
        assert truncated == c["truncated"], c
        assert sorted(sub.nodes) == sorted(ids[i] for i in c["nodes"]), c
        assert depth_by == {ids[k]: v for k, v in c["depth"]}, c
        got_edges = sorted((e.source, e.target, REL_CODE.get(enum_value(e.relationship), 31)) for e in sub.edges)
        assert got_edges == sorted((ids[s], ids[t], r) for s, t, r in c["edges"]), c


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_compute_dependency_reach(name):
    from agent_bom_b200.graph import compute_dependency_reach

    doc, g = load(name), graph(name)
    ids = doc["node_ids"]
    rep = compute_dependency_reach(g)
    want = doc["cases"]["dependency_reach"]
    assert [(p.package_id, list(p.reachable_from), p.min_hop_distance) for p in rep.packages.values()] == \
        [(ids[p], [ids[a] for a in ags], mh) for p, ags, mh in want["packages"]]
    assert [(v.vulnerability_id, list(v.package_ids), list(v.reachable_from), v.min_hop_distance) for v in rep.vulnerabilities.values()] == \
        [(ids[v], [ids[p] for p in pk], [ids[a] for a in ags], mh) for v, pk, ags, mh in want["vulnerabilities"]]


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_derived_attack_paths(name):
    from agent_bom_b200.graph import derived_attack_paths
    from agent_bom_b200.graph.schema import REL_CODE

    doc, g = load(name), graph(name)
    ids = doc["node_ids"]
    got = derived_attack_paths(g)
    want = doc["cases"]["derived_paths"]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert (a.source, a.target, a.hops) == (ids[b["source"]], ids[b["target"]], [ids[h] for h in b["hops"]])
        assert [REL_CODE[e] for e in a.edges] == b["edges"] and a.composite_risk == b["risk"]
        assert a.credential_exposure == b["creds"] and a.tool_exposure == b["tools"] and a.vuln_ids == b["vuln_ids"]


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_ranked_attack_paths_pages(name):
    """Device-side ranking: every page equals the corresponding slice of the reference's sorted AttackPath list."""
    from agent_bom_b200.graph import ranked_attack_paths
    from agent_bom_b200.graph.schema import REL_CODE

    doc, g = load(name), graph(name)
    ids = doc["node_ids"]
    want = doc["cases"]["derived_paths"]
    total_want = len(want)
    for offset, limit in ((0, total_want + 5), (0, 7), (3, 11), (max(0, total_want - 4), 10), (total_want + 3, 5)):
        got, total = ranked_attack_paths(g, offset, limit)
        assert total == total_want
        ref = want[offset: offset + limit]
        assert len(got) == len(ref)
        for a, b in zip(got, ref):
            assert (a.source, a.target, a.hops) == (ids[b["source"]], ids[b["target"]], [ids[h] for h in b["hops"]])
            assert [REL_CODE[e] for e in a.edges] == b["edges"] and a.composite_risk == b["risk"]
            assert a.credential_exposure == b["creds"] and a.tool_exposure == b["tools"]


def test_store_drop_in():
    """GraphStoreProtocol traversal subset on the GPU against the reference's answers for kat_derived (the method-by-method replay of
    the reference's own store scenarios is tests/test_gpu_store_contract.py); batched extensions; None / [] conventions."""
    from agent_bom_b200.store import B200GraphStore

    doc = load("kat_derived")
    ids = doc["node_ids"]
    g = graph_from_fixture(doc)
    g.tenant_id = "t"
    store = B200GraphStore()
    store.save_graph(g)
    assert store.latest_snapshot_id(tenant_id="t") == "golden"
    paths, reachable = store.bfs_paths(tenant_id="t", scan_id="golden", source="agent:a", max_depth=4)
    assert ["agent:a", "server:a:fs", "pkg:npm:form-data", "vuln:cve"] in paths          # reference tests/test_graph_api.py:1585-1633
    assert reachable == {p[-1] for p in paths}
    want_imp = next(c for c in doc["cases"]["impact"] if ids[c["s"]] == "vuln:cve" and c["d"] == 4)
    got_imp = store.impact_of(tenant_id="t", node_id="vuln:cve")
    assert got_imp["affected_nodes"] == sorted(ids[i] for i in want_imp["nodes"]) and got_imp["max_depth_reached"] == want_imp["maxd"]
    assert got_imp["affected_by_type"] == want_imp["by_type"] and got_imp["affected_count"] == want_imp["count"]
    assert store.impact_of(tenant_id="t", node_id="nope") is None
    sub, depth, trunc = store.traverse_subgraph(tenant_id="t", roots=["agent:a"], direction="both", max_depth=2)
    assert depth["agent:a"] == 0 and "server:a:fs" in sub.nodes and trunc is False
    sid, created, page, total = store.attack_paths(tenant_id="t", limit=2)
    want = doc["cases"]["derived_paths"]
    assert sid == "golden" and total == len(want) and len(page) == 2
    assert [p.hops for p in page] == [[ids[h] for h in w["hops"]] for w in want[:2]] and [p.composite_risk for p in page] == [w["risk"] for w in want[:2]]
    assert store.attack_paths_for_sources(tenant_id="t", source_ids={"user:u"}) == []      # persisted rows only (api/graph_store.py:792-834); none here
    many = store.impact_of_many(tenant_id="t", node_ids=["vuln:cve", "nope", "mis:m"])
    assert many[1] is None and many[0] == got_imp
    rep = store.dependency_reach(tenant_id="t")
    assert rep.packages["pkg:npm:form-data"].reachable


def test_backend_drop_in():
    """GraphBackend protocol shapes (graph_backend.py:23-38); the full behaviour is pinned in tests/test_backend_centrality.py."""
    from agent_bom_b200.backend import get_backend

    b = get_backend("b200")
    for n in "aspvx":
        b.add_node(n, "agent", n.upper())
    for s, t in (("a", "s"), ("s", "p"), ("p", "v"), ("a", "x")):
        b.add_edge(s, t, "uses", directed=True)
    assert b.node_count() == 5 and b.edge_count() == 4 and b.has_edge("a", "s") and not b.has_edge("s", "a")
    assert b.bfs("a", max_depth=2) == [["a", "s"], ["a", "x"], ["a", "s", "p"]]
    assert b.shortest_path("a", "v") == ["a", "s", "p", "v"] and b.shortest_path("v", "a") is None
    b.add_edge("v", "a", "uses")                                        # undirected by default
    assert b.shortest_path("v", "a") == ["v", "a"] and b.shortest_path("a", "v") == ["a", "v"]
    with pytest.raises(ValueError):
        get_backend("networkx")


def test_mutation_invalidates_device_cache():
    from agent_bom_b200.graph import EntityType, RelationshipType, UnifiedEdge, UnifiedGraph, UnifiedNode

    g = UnifiedGraph()
    for n, t in (("a", EntityType.AGENT), ("s", EntityType.SERVER), ("v", EntityType.VULNERABILITY)):
        g.add_node(UnifiedNode(id=n, entity_type=t, label=n))
    g.add_edge(UnifiedEdge(source="a", target="s", relationship=RelationshipType.USES))
    assert g.impact_of("v")["affected_count"] == 0
    g.add_edge(UnifiedEdge(source="s", target="v", relationship=RelationshipType.VULNERABLE_TO))
    assert g.impact_of("v")["affected_nodes"] == ["a", "s"]
    assert g.reachable_from("v") == {"v"}                                                    # tests/test_graph_schema.py:926-946


@pytest.mark.parametrize("tenant,scan", [("default", "scan-a"), ("default", "scan-b"), ("acme", "")])
def test_snapshot_graph_answers_like_a_record_graph(tmp_path, tenant, scan):
    """A topology-first snapshot (graph/snapshot.py, thin lazy records) gives the answers of a graph made of full records with the
    reference's load_graph content (tests/golden/snapshot, written and read back by the reference), and of the CPU oracle."""
    import gzip
    import json
    from pathlib import Path

    import numpy as np

    from agent_bom_b200.graph import EntityType, RelationshipType, UnifiedEdge, UnifiedGraph, UnifiedNode
    from agent_bom_b200.graph.exposure import derived_attack_paths, ranked_attack_paths
    from agent_bom_b200.graph.schema import ENTITY_CODE, REL_CODE
    from agent_bom_b200.graph.snapshot import load_snapshot
    from agent_bom_b200.store import B200GraphStore
    from oracle import oracle as orc

    gold = Path(__file__).parent / "golden" / "snapshot"
    db = tmp_path / "graph.sqlite"
    db.write_bytes(gzip.decompress((gold / "graph.sqlite.gz").read_bytes()))
    case = next(c for c in json.loads(gzip.decompress((gold / "expected.json.gz").read_bytes())) if c["tenant"] == tenant and c["scan"] == scan)
    snap = load_snapshot(db, tenant_id=tenant, scan_id=scan)
    rec = UnifiedGraph(scan_id=case["scan_id"], tenant_id=case["tenant_id"], created_at=case["created_at"])
    for nid, kind, label, sev, risk in case["nodes"]:
        rec.add_node(UnifiedNode(id=nid, entity_type=EntityType(kind), label=label, severity=sev, risk_score=risk))
    for s, t, r, direction, trav, w in case["edges"]:
        rec.add_edge(UnifiedEdge(source=s, target=t, relationship=RelationshipType(r), direction=direction, traversable=trav, weight=w))
    ids = [n[0] for n in case["nodes"]]
    assert snap.impact_of_many(ids, 4) == rec.impact_of_many(ids, 4)
    some = ids[:: max(1, len(ids) // 40)]
    assert snap.bfs_many(some, 4, True) == rec.bfs_many(some, 4, True)
    assert snap.bfs_many(some, 3, False) == rec.bfs_many(some, 3, False)
    # the oracle on the same arrays (parity anchor, not just self-consistency)
    idx = {nid: i for i, nid in enumerate(ids)}
    og = orc.build_csr(len(ids), np.asarray([idx[e[0]] for e in case["edges"]], np.int32), np.asarray([idx[e[1]] for e in case["edges"]], np.int32),
                       np.asarray([REL_CODE[e[2]] for e in case["edges"]], np.uint8),
                       np.asarray([(1 if e[4] else 0) | (2 if e[3] == "bidirectional" else 0) for e in case["edges"]], np.uint8),
                       np.asarray([ENTITY_CODE[n[1]] for n in case["nodes"]], np.uint8))
    want = orc.impact_many(og, np.arange(len(ids), dtype=np.int32), 4)
    got = snap.impact_of_many(ids, 4)
    for q in range(len(ids)):
        a, b = int(want.off[q]), int(want.off[q + 1])
        assert got[q]["affected_nodes"] == sorted(ids[i] for i in want.nodes[a:b].tolist()) and got[q]["max_depth_reached"] == int(want.maxd[q])
    for direction in ("forward", "reverse", "both"):
        for roots in (some[:1], some[1:4]):
            a_sub, a_depth, a_tr = snap.traverse_subgraph(roots, direction=direction, max_depth=3, max_nodes=60, max_edges=200)
            b_sub, b_depth, b_tr = rec.traverse_subgraph(roots, direction=direction, max_depth=3, max_nodes=60, max_edges=200)
            assert (sorted(a_sub.nodes), a_depth, a_tr) == (sorted(b_sub.nodes), b_depth, b_tr)
            assert [(e.source, e.target, e.relationship, e.direction, e.traversable) for e in a_sub.edges] == \
                   [(e.source, e.target, e.relationship, e.direction, e.traversable) for e in b_sub.edges]
            assert all(n._full is not None for n in a_sub.nodes.values())           # hydrated in bulk, ready for to_dict()
            assert all(e.to_dict()["id"] == e.id for e in a_sub.edges)
    snap_paths, rec_paths = derived_attack_paths(snap), derived_attack_paths(rec)
    if case["attack_paths"]:
        assert [p.to_dict() for p in snap_paths] == case["attack_paths"]          # materialised rows win
    else:
        assert [p.to_dict() for p in snap_paths] == [p.to_dict() for p in rec_paths]
        page, total = ranked_attack_paths(snap, 0, 25)
        assert total == len(rec_paths) and [p.to_dict() for p in page] == [p.to_dict() for p in rec_paths[:25]]

    class Inner:
        _db_path = db

        def load_graph(self, **_kw):
            raise AssertionError("the topology-first path should have served this")

    store = B200GraphStore(inner=Inner())
    probe = ids[len(ids) // 2]
    assert store.impact_of(tenant_id=tenant, scan_id=scan, node_id=probe) == rec.impact_of(probe)
    paths, reach = store.bfs_paths(tenant_id=tenant, scan_id=scan, source=some[0], max_depth=4)
    assert paths == rec.bfs(some[0], 4, True)

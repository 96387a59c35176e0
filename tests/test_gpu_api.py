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
    """GraphStoreProtocol traversal subset on the GPU: same answers as the engine-level calls; None / [] conventions."""
    from agent_bom_b200.store import B200GraphStore

    doc = load("kat_derived")
    g = graph_from_fixture(doc)
    store = B200GraphStore()
    store.save_graph(g)
    assert store.latest_snapshot_id(tenant_id="t") == "golden"
    paths, reachable = store.bfs_paths(tenant_id="t", scan_id="golden", source="agent:a", max_depth=4)
    assert paths == g.bfs("agent:a", 4, True) and reachable == g.reachable_from("agent:a", 4, traversable_only=True, include_source=False)
    assert ["agent:a", "server:a:fs", "pkg:npm:form-data", "vuln:cve"] in paths          # reference tests/test_graph_api.py:1585-1633
    assert store.impact_of(tenant_id="t", node_id="vuln:cve") == g.impact_of("vuln:cve")
    assert store.impact_of(tenant_id="t", node_id="nope") is None
    sub, depth, trunc = store.traverse_subgraph(tenant_id="t", roots=["agent:a"], direction="both", max_depth=2)
    assert depth["agent:a"] == 0 and "server:a:fs" in sub.nodes and trunc is False
    sid, created, page, total = store.attack_paths(tenant_id="t", limit=2)
    assert sid == "golden" and total == len(doc["cases"]["derived_paths"]) and len(page) == 2
    assert page[0].composite_risk >= page[1].composite_risk
    want = doc["cases"]["derived_paths"]
    assert [p.hops for p in page] == [[doc["node_ids"][h] for h in w["hops"]] for w in want[:2]]
    only = store.attack_paths_for_sources(tenant_id="t", source_ids={"user:u"})
    assert only and all(p.source == "user:u" for p in only)
    many = store.impact_of_many(tenant_id="t", node_ids=["vuln:cve", "nope", "mis:m"])
    assert many[1] is None and many[0] == g.impact_of("vuln:cve")
    rep = store.dependency_reach(tenant_id="t")
    assert rep.packages["pkg:npm:form-data"].reachable


def test_backend_drop_in():
    from agent_bom_b200.backend import get_backend

    b = get_backend("b200")
    for s, t in (("a", "s"), ("s", "p"), ("p", "v"), ("a", "x")):
        b.add_edge(s, t)
    assert b.node_count() == 5 and b.edge_count() == 4 and b.has_edge("a", "s") and not b.has_edge("s", "a")
    assert b.bfs("a", max_depth=2) == ["s", "x", "p"]
    assert b.shortest_path("a", "v") == ["a", "s", "p", "v"] and b.shortest_path("v", "a") is None
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

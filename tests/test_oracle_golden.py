"""Pin the CPU oracle (oracle/oracle.c) against outputs of the unmodified reference engine.

The fixtures under tests/golden/ were produced by oracle/make_golden.py, which
imports /root/reference/src and rebuilds the reference's own known-answer
graphs (tests/test_graph_schema.py:328-441,895-946; tests/test_dependency_reach.py:26-161;
tests/test_graph_api.py:1585-1633) plus generator estates.  CPU-only.
"""

from __future__ import annotations

import pytest

from golden_util import ALL_FIXTURES, SMALL_FIXTURES, load, node_rank, oracle_graph, rank_path_rows
from oracle import oracle as orc

ENTITY_VALUES = (
    "agent server package tool model dataset container cloud_resource vulnerability misconfiguration credential "
    "org account user group role policy service_account service_principal federated_identity provider environment fleet cluster"
).split()
DYN = (1 << 26) | (1 << 27) | (1 << 28)


def paths_from_parents(src: int, nodes, parent):
    paths = []
    for i in range(len(nodes)):
        p = parent[i]
        paths.append(([src] if p < 0 else paths[p]) + [int(nodes[i])])
    return paths


@pytest.mark.parametrize("name", SMALL_FIXTURES)
def test_adjacency_model(name):
    """a1: CSR rows == adjacency / reverse_adjacency lists in list order (container.py:146-198)."""
    doc, g = load(name), oracle_graph(name)
    for key, off, nbr, meta in (("adjacency", g.fwd_off, g.fwd_nbr, g.fwd_meta), ("reverse_adjacency", g.rev_off, g.rev_nbr, g.rev_meta)):
        want = {int(k): v for k, v in doc[key].items()}
        for u in range(g.n_nodes):
            got = [[int(nbr[p]), int(meta[p]) & 0x1F] for p in range(int(off[u]), int(off[u + 1]))]
            assert got == want.get(u, []), (key, u)


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_impact_of(name):
    doc, g = load(name), oracle_graph(name)
    cases = doc["cases"]["impact"]
    for depth in sorted({c["d"] for c in cases}):
        sub = [c for c in cases if c["d"] == depth]
        res = orc.impact_many(g, [c["s"] for c in sub], depth)
        for q, c in enumerate(sub):
            nodes, _ = res.slice(q)
            assert sorted(int(x) for x in nodes) == c["nodes"]
            assert len(nodes) == c["count"]
            assert int(res.maxd[q]) == c["maxd"]
            hist = {ENTITY_VALUES[t]: int(n) for t, n in enumerate(res.hist[q]) if n}
            assert hist == c["by_type"]
    # missing / ghost source -> zeros (container.py:247-248)
    res = orc.impact_many(g, [-1, g.n_nodes + 5], 4)
    assert res.off[-1] == 0 and list(res.flags) == [2, 2]


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_bfs_paths_exact_order(name):
    doc, g = load(name), oracle_graph(name)
    cases = doc["cases"]["bfs"]
    for key in sorted({(c["d"], c["t"]) for c in cases}):
        sub = [c for c in cases if (c["d"], c["t"]) == key]
        res = orc.bfs_many(g, [c["s"] for c in sub], key[0], key[1])
        for q, c in enumerate(sub):
            nodes, parent = res.slice(q)
            assert paths_from_parents(c["s"], nodes, parent) == c["paths"]


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_reachable_from(name):
    doc, g = load(name), oracle_graph(name)
    cases = doc["cases"]["reachable"]
    for key in sorted({(c["d"], c["t"]) for c in cases}):
        sub = [c for c in cases if (c["d"], c["t"]) == key]
        res = orc.reachable_many(g, [c["s"] for c in sub], key[0], key[1])
        for q, c in enumerate(sub):
            nodes, _ = res.slice(q)
            got = {int(x) for x in nodes} | ({c["s"]} if c["inc"] else set())
            assert sorted(got) == c["nodes"]


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_shortest_path(name):
    doc, g = load(name), oracle_graph(name)
    for c in doc["cases"]["shortest"]:
        got = orc.shortest_path(g, c["a"], c["b"])
        assert (None if got is None else [int(x) for x in got]) == c["path"]


def traverse_expect(g, res, q, roots, include_roots, n_real):
    """Host-side reconstruction of (sub.nodes, depth_by_node, sub.edges) from a walk slice (container.py:525-537)."""
    nodes, depth = res.slice(q)
    valid = [r for r in roots if 0 <= r < n_real]
    nr = len(valid)
    depth_by = {}
    for u, d in zip(nodes, depth):
        depth_by[int(u)] = int(d)
    visited = {int(u) for u in nodes[nr:]}
    if include_roots:
        visited |= set(valid)
    sub_nodes = {u for u in visited if g.node_type[u] != 255}
    a, b = int(res.eoff[q]), int(res.eoff[q + 1])
    edges = set()
    for eid2 in res.edges[a:b]:
        eid2 = int(eid2)
        # locate (src,dst,rel) of the original; reversed copy swaps endpoints
        s, t, r = g._edge_lookup[eid2 >> 1]
        if eid2 & 1:
            s, t = t, s
        if s in sub_nodes and t in sub_nodes:
            edges.add((s, t, r))
    return sorted(sub_nodes), sorted([k, v] for k, v in depth_by.items()), sorted(list(e) for e in edges)


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_traverse_subgraph(name):
    doc, g = load(name), oracle_graph(name)
    g._edge_lookup = [(e[0], e[1], e[2]) for e in doc["edges"]]
    n_real = doc["n_real"]
    for c in doc["cases"]["traverse"]:
        kw = c["kw"]
        mask = 0xFFFFFFFF
        if kw.get("relationship_types"):
            mask = 0
            for r in kw["relationship_types"]:
                mask |= 1 << r
        if kw.get("static_only"):
            mask &= ~DYN
        if kw.get("dynamic_only"):
            mask &= DYN
        roots = c["roots"]
        direction = {"forward": 1, "reverse": 2, "both": 3}[c["direction"]]
        inc = kw.get("include_roots", True)
        res = orc.traverse_many(
            g, roots, [0, len(roots)], direction=direction, max_depth=kw.get("max_depth", 4), max_nodes=kw.get("max_nodes", 500),
            max_edges=kw.get("max_edges", 10_000), rel_mask=mask, traversable_only=kw.get("traversable_only", False), include_roots=inc,
        )
        nodes, depth, edges = traverse_expect(g, res, 0, roots, inc, n_real)
        assert bool(res.flags[0] & 1) == c["truncated"], c
        assert nodes == c["nodes"], c
        assert depth == c["depth"], c
        assert edges == c["edges"], c


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_dependency_reach(name):
    doc, g = load(name), oracle_graph(name)
    rank = node_rank(doc["node_ids"])
    from agent_bom_b200.graph.schema import REACH_MASK, VULN_PKG_MASK

    out = orc.dependency_reach(g, doc["agents"], REACH_MASK, VULN_PKG_MASK, rank)
    want = doc["cases"]["dependency_reach"]
    got_pk = [[int(p), [int(a) for a in out["pkg_agents"][out["pkg_off"][i]: out["pkg_off"][i + 1]]], int(out["pkg_minhop"][i])]
              for i, p in enumerate(out["pkg_ids"])]
    assert got_pk == want["packages"]
    got_v = []
    for i, v in enumerate(out["vuln_ids"]):
        got_v.append([int(v), [int(x) for x in out["vuln_pkgs"][out["vuln_poff"][i]: out["vuln_poff"][i + 1]]],
                      [int(x) for x in out["vuln_agents"][out["vuln_aoff"][i]: out["vuln_aoff"][i + 1]]], int(out["vuln_minhop"][i])])
    assert got_v == want["vulnerabilities"]


@pytest.mark.parametrize("name", [n for n in ALL_FIXTURES])
def test_derived_attack_paths(name):
    doc, g = load(name), oracle_graph(name)
    if "derived_paths" not in doc["cases"]:
        pytest.skip("no derived paths in fixture")
    rank = node_rank(doc["node_ids"])
    rows = orc.derived_paths(g, doc["findings"], rank)
    got = rank_path_rows(doc, rows, g)
    want = doc["cases"]["derived_paths"]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a == b

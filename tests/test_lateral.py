"""Lateral-path search: oracle vs the reference's find_lateral_paths (CPU), CUDA search vs both (GPU).

``tests/golden/context/lateral.json.gz`` (``oracle/make_golden.py --lateral-only``): ContextGraphs built by the unmodified
reference and its ``find_lateral_paths`` answers for agent / server / tool / credential / unknown sources at several depths —
including parallel edges (the same node sequence arriving twice), ids without a node record and a node with the source
agent's own label.
"""

from __future__ import annotations

import gzip
import json
from dataclasses import asdict
from pathlib import Path

import pytest

from agent_bom_b200.context_graph import ContextGraph, EdgeKind, GraphEdge, GraphNode, NodeKind

DOCS = json.loads(gzip.decompress((Path(__file__).parent / "golden" / "context" / "lateral.json.gz").read_bytes()))
IDS = [d["name"] for d in DOCS]


def rebuild(doc) -> ContextGraph:
    g = ContextGraph()
    for nid, kind, label, meta in doc["nodes"]:
        g.add_node(GraphNode(id=nid, kind=NodeKind(kind), label=label, metadata=dict(meta)))
    for s, t, k, meta in doc["edges"]:
        g.add_edge(GraphEdge(source=s, target=t, kind=EdgeKind(k), metadata=dict(meta)))
    assert {nid: [[e.source, e.target, e.kind.value] for e in lst] for nid, lst in g.adjacency.items() if lst} == doc["adjacency"]
    return g


def as_dicts(paths):
    out = []
    for p in paths:
        d = asdict(p)
        d["edges"] = [getattr(k, "value", k) for k in p.edges]
        out.append(d)
    return out


@pytest.mark.parametrize("doc", DOCS, ids=IDS)
def test_oracle_matches_the_reference(doc):
    from oracle import lateral_oracle as lo

    g = rebuild(doc)
    for c in doc["cases"]:
        assert lo.find(g, c["source"], c["max_depth"]) == c["paths"], (c["source"], c["max_depth"])


def test_path_scoring_on_the_host():
    """build_lateral_path alone (no device): the reference's paths, re-scored from their hops and edge kinds."""
    from agent_bom_b200.lateral import build_lateral_path

    for doc in DOCS:
        g = rebuild(doc)
        for c in doc["cases"][:: max(1, len(doc["cases"]) // 12)]:
            for want in c["paths"]:
                got = build_lateral_path(g, want["source"], want["target"], list(want["hops"]), [EdgeKind(k) for k in want["edges"]])
                assert as_dicts([got])[0] == want


@pytest.mark.gpu
@pytest.mark.parametrize("doc", DOCS, ids=IDS)
def test_device_search_matches_the_reference(doc):
    from agent_bom_b200.lateral import find_lateral_paths, find_lateral_paths_many

    g = rebuild(doc)
    for depth in sorted({c["max_depth"] for c in doc["cases"]}):
        cases = [c for c in doc["cases"] if c["max_depth"] == depth]
        got = find_lateral_paths_many(g, [c["source"] for c in cases], depth)          # one launch for all sources of this depth
        for c, paths in zip(cases, got):
            assert as_dicts(paths) == c["paths"], (c["source"], depth)
    c = doc["cases"][0]
    assert as_dicts(find_lateral_paths(g, c["source"], c["max_depth"])) == c["paths"]
    assert find_lateral_paths(g, "no-such-node") == []


@pytest.mark.gpu
def test_device_search_against_the_oracle_where_the_caps_bite():
    """A dense fleet: every search hits the 100-path cap, and the 10 000-queue cap throttles expansion; plus a hub whose
    adjacency row is longer than a warp."""
    import random

    from agent_bom_b200.lateral import search_many
    from oracle import lateral_oracle as lo

    rng = random.Random(3)
    g = ContextGraph()
    for a in range(60):
        g.add_node(GraphNode(id=f"agent:{a}", kind=NodeKind.AGENT, label=f"a{a}"))
    for s in range(40):
        owner = rng.randrange(60)
        g.add_node(GraphNode(id=f"server:{s}", kind=NodeKind.SERVER, label=f"s{s % 9}", metadata={"agent": f"a{owner}"}))
        for a in rng.sample(range(60), rng.randint(2, 9)):
            g.add_edge(GraphEdge(source=f"agent:{a}", target=f"server:{s}", kind=EdgeKind.USES))
        for t in range(rng.randint(1, 4)):
            tid = f"tool:{s}:{t}"
            g.add_node(GraphNode(id=tid, kind=NodeKind.TOOL, label=f"t{t}", metadata={"agent": f"a{owner}" if rng.random() < 0.7 else "", "capabilities": ["execute"]}))
            g.add_edge(GraphEdge(source=f"server:{s}", target=tid, kind=EdgeKind.PROVIDES))
    for _ in range(150):
        a, b = rng.sample(range(60), 2)
        g.add_edge(GraphEdge(source=f"agent:{a}", target=f"agent:{b}", kind=rng.choice([EdgeKind.SHARES_SERVER, EdgeKind.SHARES_CREDENTIAL]),
                             metadata={"server": f"s{rng.randrange(9)}", "credential": "K"}))
    g.add_node(GraphNode(id="server:hub", kind=NodeKind.SERVER, label="hub"))
    for a in range(60):
        g.add_edge(GraphEdge(source=f"agent:{a}", target="server:hub", kind=EdgeKind.USES))
    sources = [f"agent:{a}" for a in range(0, 60, 7)] + ["server:hub", "server:3", "tool:5:0"]
    for depth in (2, 4, 6):
        got, ms = search_many(g, sources, depth)
        for s, found in zip(sources, got):
            want = lo.search(g, s, depth)
            assert [(h, [k.value for k in ks]) for h, ks in found] == [(h, [k.value for k in ks]) for h, ks in want], (s, depth)
        assert ms >= 0.0
    # a graph where nothing is lateral: the search drains the whole depth-limited path tree and records nothing
    lone = ContextGraph()
    lone.add_node(GraphNode(id="agent:x", kind=NodeKind.AGENT, label="x"))
    for i in range(40):
        lone.add_node(GraphNode(id=f"server:{i}", kind=NodeKind.SERVER, label=f"s{i}", metadata={"agent": "x"}))
        lone.add_edge(GraphEdge(source="agent:x", target=f"server:{i}", kind=EdgeKind.USES))
        if i:
            lone.add_edge(GraphEdge(source=f"server:{i - 1}", target=f"server:{i}", kind=EdgeKind.USES))
    got, _ = search_many(lone, ["agent:x", "server:7"], 5)
    assert got == [[], []] and lo.search(lone, "agent:x", 5) == []
    with pytest.raises(Exception):
        search_many(lone, ["agent:x"], 9)                                    # max_depth beyond the record width is refused, not truncated


def test_adjacency_encoding_is_cached_per_graph_and_rebuilt_on_change():
    """The array encoding of ``graph.adjacency`` (the >99 % of a call that was Python) is reused while the graph is unchanged."""
    from types import SimpleNamespace as NS

    from agent_bom_b200 import lateral

    class G:
        pass

    g = G()
    g.nodes = {f"agent:{i}": NS(kind="agent", label=f"a{i}", metadata={}) for i in range(4)}
    g.edges = []
    g.adjacency = {"agent:0": [NS(target="agent:1", kind="shares_server")], "agent:1": [NS(target="agent:2", kind="shares_server")]}
    a1 = lateral._arrays_for(g)
    assert lateral._arrays_for(g) is a1
    g.adjacency["agent:2"] = [NS(target="agent:3", kind="shares_credential")]
    a2 = lateral._arrays_for(g)
    assert a2 is not a1 and len(a2.nbr) == 3 and lateral._arrays_for(g) is a2
    key = id(g)
    del g, a1, a2
    import gc

    gc.collect()
    assert key not in lateral._ARRAYS_CACHE

"""GraphBackend drop-in and centrality (SURVEY §8 b / f4) against the reference's InMemoryBackend and UnifiedGraph.

``tests/golden/context/centrality.json.gz`` (``oracle/make_golden.py --centrality-only``): op sequences replayed on the
reference's ``InMemoryBackend`` with every protocol method's answer, the reference's ``from_unified_graph`` bridge on the
mesh inventory, and ``UnifiedGraph.degree_centrality`` / ``bottleneck_nodes`` on the walk-fixture graphs.
"""

from __future__ import annotations

import gzip
import json
from pathlib import Path

import pytest

from golden_util import graph_from_fixture, load

GOLD = json.loads(gzip.decompress((Path(__file__).parent / "golden" / "context" / "centrality.json.gz").read_bytes()))
BACKENDS = GOLD["backends"]
LABELS = [d["label"] for d in BACKENDS]


def replay(doc):
    from agent_bom_b200.backend import get_backend

    b = get_backend("b200")
    for op in doc["ops"]:
        if op[0] == "n":
            b.add_node(op[1], op[2], op[3], **op[4])
        else:
            b.add_edge(op[1], op[2], op[3], op[4], directed=op[5], **op[6])
    return b


def pairs(result):
    return [[nid, score] for nid, score in result]


@pytest.mark.parametrize("doc", BACKENDS, ids=LABELS)
def test_backend_host_methods(doc):
    b = replay(doc)
    assert b.node_count() == doc["node_count"] and b.edge_count() == doc["edge_count"]
    assert b.to_dict() == doc["to_dict"]
    assert b.centrality_scores() == doc["centrality"]["value"]
    for nid, want in doc["neighbors"].items():
        assert b.neighbors(nid) == want
    for s, t, want in doc["has_edge"]:
        assert b.has_edge(s, t) is want
    assert b.has_node("nope") is False and b.neighbors("nope") == []


def test_backend_protocol_shape():
    """Same call shapes as the reference protocol (graph_backend.py:23-38); other backend names are an error, not a fallback."""
    import inspect

    from agent_bom_b200.backend import B200Backend, get_backend

    sig = inspect.signature(B200Backend.add_edge)
    assert list(sig.parameters)[:5] == ["self", "source", "target", "kind", "weight"] and sig.parameters["directed"].kind is inspect.Parameter.KEYWORD_ONLY
    assert list(inspect.signature(B200Backend.add_node).parameters)[:4] == ["self", "node_id", "kind", "label"]
    for name in ("has_node", "has_edge", "neighbors", "bfs", "shortest_path", "node_count", "edge_count", "to_dict", "centrality_scores", "bottleneck_nodes"):
        assert callable(getattr(B200Backend, name))
    with pytest.raises(ValueError):
        get_backend("networkx")


@pytest.mark.parametrize("name", sorted(GOLD["unified"]))
def test_degree_centrality_of_fixture_graphs(name):
    g = graph_from_fixture(load(name))
    want = GOLD["unified"][name]
    assert len(g.nodes) == want["n_nodes"]
    assert g.degree_centrality() == want["degree"]["value"]


def check_bridge_host_side(b, want):
    def no_weight(edges):          # the walk fixtures do not carry edge weights; everything else must agree
        return [{k: v for k, v in e.items() if k != "weight"} for e in edges]

    assert (b.node_count(), b.edge_count()) == (want["n_nodes"], want["n_edges"])
    assert no_weight(b.to_dict()["edges"]) == no_weight(want["to_dict"]["edges"]) and b.to_dict()["stats"] == want["to_dict"]["stats"]
    assert [n["id"] for n in b.to_dict()["nodes"]] == [n["id"] for n in want["to_dict"]["nodes"]]
    assert b.centrality_scores() == want["centrality"]


def test_unified_graph_bridge_host_side():
    from agent_bom_b200.backend import from_unified_graph

    check_bridge_host_side(from_unified_graph(graph_from_fixture(load("mesh_inventory"))), GOLD["mesh_bridge"])


# ── GPU ─────────────────────────────────────────────────────────────────────

@pytest.mark.gpu
@pytest.mark.parametrize("doc", BACKENDS, ids=LABELS)
def test_backend_traversals(doc):
    b = replay(doc)
    for s, d, want in doc["bfs"]:
        assert b.bfs(s, d) == want, (s, d)
    for s, t, want in doc["shortest"]:
        assert b.shortest_path(s, t) == want, (s, t)
    for key, top in (("bottleneck_5", 5), ("bottleneck_20", 20)):
        want = doc[key]
        if "key_error" in want:
            with pytest.raises(KeyError):
                b.bottleneck_nodes(top_n=top)
        else:
            assert pairs(b.bottleneck_nodes(top_n=top)) == want["value"], key
    # a mutation after the first query re-exports the adjacency
    b.add_node("late", "agent", "late")
    b.add_edge("late", doc["ops"][0][1] if doc["ops"] else "late", "uses", directed=True)
    assert b.bfs("late", 1) == ([["late", doc["ops"][0][1]]] if doc["ops"] else [])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD["unified"]))
def test_bottleneck_nodes_of_fixture_graphs(name):
    g = graph_from_fixture(load(name))
    want = GOLD["unified"][name]
    for key, top in (("bottleneck_5", 5), ("bottleneck_12", 12)):
        if "key_error" in want[key]:
            with pytest.raises(KeyError):
                g.bottleneck_nodes(top_n=top)
        else:
            assert pairs(g.bottleneck_nodes(top_n=top)) == want[key]["value"], key


@pytest.mark.gpu
def test_unified_graph_bridge_matches_the_reference_bridge():
    from agent_bom_b200.backend import from_context_graph, from_unified_graph

    g = graph_from_fixture(load("mesh_inventory"))
    want = GOLD["mesh_bridge"]
    b = from_unified_graph(g)
    check_bridge_host_side(b, want)
    assert pairs(b.bottleneck_nodes(top_n=10)) == want["bottleneck_10"]["value"]
    for s, d, paths in want["bfs"]:
        assert b.bfs(s, d) == paths
    c = from_context_graph({"nodes": [{"id": "a", "kind": "agent", "label": "A"}, {"id": "s", "kind": "server"}], "edges": [{"source": "a", "target": "s", "kind": "uses"}]})
    assert c.bfs("s", 2) == [["s", "a"]] and c.shortest_path("a", "s") == ["a", "s"] and c.edge_count() == 1

"""The reference's own backend tests (``/root/reference/tests/test_graph_backend.py``: TestInMemoryBackend :56-149, TestProtocol
:244-247, TestFactory :262-268, TestFromContextGraph :288-330), re-stated for ``B200Backend`` — same graphs (triangle, chain, star),
same assertions.  Host-only checks run everywhere; the traversals need the GPU."""

from __future__ import annotations

import pytest

from agent_bom_b200.backend import B200Backend, GraphBackend, from_context_graph, get_backend


def build_triangle(b):
    b.add_node("A", kind="agent", label="Agent A")
    b.add_node("B", kind="server", label="Server B")
    b.add_node("C", kind="credential", label="Cred C")
    b.add_edge("A", "B", kind="uses", weight=1.0)
    b.add_edge("B", "C", kind="exposes", weight=2.0)
    b.add_edge("C", "A", kind="shares", weight=1.5)
    return b


def build_chain(b):
    for name in "ABCDE":
        b.add_node(name, kind="node", label=f"Node {name}")
    for s, t in ("AB", "BC", "CD", "DE"):
        b.add_edge(s, t, kind="link")
    return b


def build_star(b):
    b.add_node("center", kind="server", label="Central Server")
    for i in range(5):
        b.add_node(f"spoke-{i}", kind="agent", label=f"Agent {i}")
        b.add_edge("center", f"spoke-{i}", kind="uses")
    return b


def test_add_node():
    g = B200Backend()
    g.add_node("n1", kind="agent", label="Agent 1")
    assert g.has_node("n1") and not g.has_node("n2") and g.node_count() == 1


def test_add_edge_is_bidirectional_by_default():
    g = B200Backend()
    g.add_node("A", kind="agent", label="A")
    g.add_node("B", kind="server", label="B")
    g.add_edge("A", "B", kind="uses")
    assert g.has_edge("A", "B") and g.has_edge("B", "A") and g.edge_count() == 1


def test_neighbors():
    g = build_triangle(B200Backend())
    assert {"B", "C"} <= set(g.neighbors("A"))


def test_to_dict():
    data = build_triangle(B200Backend()).to_dict()
    assert len(data["nodes"]) == 3 and len(data["edges"]) == 3
    assert data["stats"] == {"node_count": 3, "edge_count": 3}


def test_centrality_scores():
    scores = build_star(B200Backend()).centrality_scores()
    assert scores["center"] == max(scores.values())
    assert B200Backend().centrality_scores() == {} and B200Backend().bottleneck_nodes() == []


def test_protocol_and_factory():
    assert isinstance(B200Backend(), GraphBackend)
    assert isinstance(get_backend("b200"), B200Backend) and isinstance(get_backend("auto"), GraphBackend)


def test_from_context_graph():
    data = {"nodes": [{"id": "agent:test", "kind": "agent", "label": "test"}, {"id": "server:test:mcp", "kind": "server", "label": "mcp"}],
            "edges": [{"source": "agent:test", "target": "server:test:mcp", "kind": "uses", "weight": 1.0}]}
    g = from_context_graph(data)
    assert g.has_node("agent:test") and g.has_node("server:test:mcp") and g.has_edge("agent:test", "server:test:mcp")
    assert (g.node_count(), g.edge_count()) == (2, 1)
    assert from_context_graph({}).node_count() == 0
    hub = from_context_graph({"nodes": [{"id": n, "kind": "agent", "label": n} for n in ("center", "a1", "a2", "a3")],
                              "edges": [{"source": "center", "target": t, "kind": "uses", "weight": 1.0} for t in ("a1", "a2", "a3")]})
    scores = hub.centrality_scores()
    assert scores["center"] == max(scores.values())


def test_nonexistent_nodes_need_no_device():
    g = B200Backend()
    assert g.bfs("nonexistent") == [] and g.shortest_path("X", "Y") is None
    chain = build_chain(B200Backend())
    assert chain.shortest_path("A", "A") == ["A"]


@pytest.mark.gpu
def test_bfs_and_depth_limit():
    g = build_chain(B200Backend())
    ends = {p[-1] for p in g.bfs("A", max_depth=4)}
    assert {"B", "E"} <= ends
    ends = {p[-1] for p in g.bfs("A", max_depth=2)}
    assert "B" in ends and "C" in ends and "D" not in ends


@pytest.mark.gpu
def test_shortest_paths():
    g = build_chain(B200Backend())
    assert g.shortest_path("A", "E") == ["A", "B", "C", "D", "E"]
    lone = B200Backend()
    lone.add_node("X", kind="node", label="X")
    lone.add_node("Y", kind="node", label="Y")
    assert lone.shortest_path("X", "Y") is None


@pytest.mark.gpu
def test_bottleneck_nodes():
    bottlenecks = build_chain(B200Backend()).bottleneck_nodes(top_n=3)
    assert len(bottlenecks) <= 3
    assert any(n in [b[0] for b in bottlenecks] for n in "BCD")
    assert bottlenecks == [("C", 8 / 20), ("B", 6 / 20), ("D", 6 / 20)]          # what the reference's InMemoryBackend returns for this chain

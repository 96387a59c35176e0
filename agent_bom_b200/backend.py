"""B200Backend — GraphBackend-protocol drop-in (reference ``/root/reference/src/agent_bom/graph_backend.py:23-38``).

The reference's analysis backends (``InMemoryBackend`` :41-155, ``NetworkXBackend``
:158-242) hold a plain directed graph of ids with attribute dicts and expose
``bfs(source, max_depth)`` / ``shortest_path``; ``get_backend(name)`` (:245-259)
picks one.  This backend keeps the same methods and runs the two traversals on
the GPU through the typed engine (every edge gets the neutral relationship
``uses``; InMemoryBackend's BFS is direction-aware and depth-limited exactly
like ``UnifiedGraph.reachable_from``).
"""

from __future__ import annotations

from typing import Any

from .graph.container import UnifiedGraph
from .graph.model import UnifiedEdge, UnifiedNode
from .graph.schema import EntityType, RelationshipType, enum_value


class B200Backend:
    name = "b200"

    def __init__(self, device: int = 0):
        self._g = UnifiedGraph(device=device)
        self._attrs: dict[str, dict[str, Any]] = {}
        self._edge_attrs: dict[tuple[str, str], dict[str, Any]] = {}

    def add_node(self, node_id: str, **attrs: Any) -> None:
        if node_id not in self._g.nodes:
            self._g.add_node(UnifiedNode(id=node_id, entity_type=EntityType.AGENT, label=str(attrs.get("label", node_id))))
        self._attrs.setdefault(node_id, {}).update(attrs)

    def add_edge(self, source: str, target: str, **attrs: Any) -> None:
        for n in (source, target):
            if n not in self._g.nodes:
                self.add_node(n)
        self._g.add_edge(UnifiedEdge(source=source, target=target, relationship=RelationshipType.USES))
        self._edge_attrs.setdefault((source, target), {}).update(attrs)

    def has_node(self, node_id: str) -> bool:
        return node_id in self._g.nodes

    def has_edge(self, source: str, target: str) -> bool:
        return (source, target) in self._edge_attrs

    def neighbors(self, node_id: str) -> list[str]:
        return self._g.neighbors(node_id)

    def bfs(self, source: str, max_depth: int = 4) -> list[str]:
        """Node ids reachable from ``source`` within ``max_depth`` hops, in discovery order (source excluded)."""
        if source not in self._g.nodes:
            return []
        g = self._g
        import numpy as np

        res = g.device_graph.reachable_many(np.asarray([g.csr.idx(source)], dtype=np.int32), max_depth, False)
        return [g.csr.node_ids[i] for i in res.slice(0).tolist()]

    def shortest_path(self, source: str, target: str) -> list[str] | None:
        return self._g.shortest_path(source, target)

    def node_count(self) -> int:
        return len(self._g.nodes)

    def edge_count(self) -> int:
        return len(self._g.edges)

    def to_dict(self) -> dict[str, Any]:
        return {"nodes": [{"id": n, **self._attrs.get(n, {})} for n in self._g.nodes],
                "edges": [{"source": s, "target": t, **a} for (s, t), a in self._edge_attrs.items()]}

    @classmethod
    def from_unified_graph(cls, graph, device: int = 0) -> "B200Backend":
        """Bridge used like the reference's ``from_unified_graph`` (graph_backend.py:281-306)."""
        b = cls(device=device)
        for n in graph.nodes.values():
            b.add_node(n.id, label=n.label, entity_type=enum_value(n.entity_type))
        for e in graph.edges:
            b.add_edge(e.source, e.target, relationship=enum_value(e.relationship))
            if getattr(e, "direction", "directed") == "bidirectional":
                b.add_edge(e.target, e.source, relationship=enum_value(e.relationship))
        return b


def get_backend(backend: str = "b200", **kwargs):
    """``get_backend("b200")`` — the name a ``--graph-backend b200`` flag would pass (reference cli/options_surfaces.py:161-162)."""
    if backend in ("b200", "auto"):
        return B200Backend(**kwargs)
    raise ValueError(f"unknown graph backend {backend!r}: this package only provides 'b200' (no CPU fallback)")

"""B200Backend — drop-in for the reference's ``GraphBackend`` protocol (``/root/reference/src/agent_bom/graph_backend.py:23-38``).

Same methods, argument meaning and return shapes as ``InMemoryBackend`` (``:41-155``): nodes are attribute dicts,
edges live in a dict-of-dicts adjacency (``add_edge`` is undirected unless ``directed=True``; re-adding a pair
overwrites its attributes but keeps its place in the neighbour order), ``bfs`` returns one path per reached node in
discovery order, ``shortest_path`` the first-discoverer path, ``centrality_scores`` degree centrality and
``bottleneck_nodes`` the 50-source sampled score.  The traversals (``bfs``, ``shortest_path``, ``bottleneck_nodes``)
run on the GPU through the ordered walk engine; the adjacency is exported to a CSR lazily and re-exported after a
mutation.  ``get_backend`` / ``from_context_graph`` / ``from_unified_graph`` mirror ``:245-306``.
"""

from __future__ import annotations

from collections import defaultdict
from typing import Protocol, runtime_checkable

import numpy as np

from . import _lib
from .graph import csr as csrmod
from .graph.schema import ENTITY_CODE_GHOST, enum_value

_SAMPLE = 50      # sources sampled by bottleneck_nodes (graph_backend.py:135)


@runtime_checkable
class GraphBackend(Protocol):
    """The reference's protocol (graph_backend.py:23-38), restated so callers can ``isinstance``-check a backend."""

    def add_node(self, node_id: str, kind: str, label: str, **metadata: object) -> None: ...
    def add_edge(self, source: str, target: str, kind: str, weight: float = 1.0, *, directed: bool = False, **metadata: object) -> None: ...
    def has_node(self, node_id: str) -> bool: ...
    def has_edge(self, source: str, target: str) -> bool: ...
    def neighbors(self, node_id: str) -> list[str]: ...
    def bfs(self, source: str, max_depth: int = 4) -> list[list[str]]: ...
    def shortest_path(self, source: str, target: str) -> list[str] | None: ...
    def node_count(self) -> int: ...
    def edge_count(self) -> int: ...
    def to_dict(self) -> dict: ...
    def centrality_scores(self) -> dict[str, float]: ...
    def bottleneck_nodes(self, top_n: int = 5) -> list[tuple[str, float]]: ...


class B200Backend:
    name = "b200"

    def __init__(self, device: int = 0):
        self._nodes: dict[str, dict] = {}
        self._adj: dict[str, dict[str, dict]] = defaultdict(dict)
        self._edge_count = 0
        self._device = device
        self._csr: csrmod.HostCSR | None = None
        self._dg = None

    # ── mutation ────────────────────────────────────────────────────────
    def _invalidate(self) -> None:
        self._csr = None
        if self._dg is not None:
            self._dg.close()
            self._dg = None

    def add_node(self, node_id: str, kind: str, label: str, **metadata: object) -> None:
        self._nodes[node_id] = {"kind": kind, "label": label, **metadata}
        self._invalidate()

    def add_edge(self, source: str, target: str, kind: str, weight: float = 1.0, *, directed: bool = False, **metadata: object) -> None:
        self._adj[source][target] = {"kind": kind, "weight": weight, **metadata}
        if not directed:
            self._adj[target][source] = {"kind": kind, "weight": weight, **metadata}
        self._edge_count += 1
        self._invalidate()

    # ── point queries (host) ────────────────────────────────────────────
    def has_node(self, node_id: str) -> bool:
        return node_id in self._nodes

    def has_edge(self, source: str, target: str) -> bool:
        return target in self._adj.get(source, {})

    def neighbors(self, node_id: str) -> list[str]:
        return list(self._adj.get(node_id, {}).keys())

    def node_count(self) -> int:
        return len(self._nodes)

    def edge_count(self) -> int:
        return self._edge_count

    def to_dict(self) -> dict:
        return {
            "nodes": [{"id": nid, **data} for nid, data in self._nodes.items()],
            "edges": [{"source": src, "target": tgt, **data} for src, targets in self._adj.items() for tgt, data in targets.items() if src < tgt],
            "stats": {"node_count": self.node_count(), "edge_count": self.edge_count()},
        }

    def centrality_scores(self) -> dict[str, float]:
        if not self._nodes:
            return {}
        max_possible = max(len(self._nodes) - 1, 1)
        return {nid: len(self._adj.get(nid, {})) / max_possible for nid in self._nodes}

    # ── device export ───────────────────────────────────────────────────
    @property
    def csr(self) -> csrmod.HostCSR:
        """Adjacency rows in neighbour (dict) order; ids without a node record become ghost rows after the real ones."""
        if self._csr is None:
            ids = list(self._nodes)
            index = {nid: i for i, nid in enumerate(ids)}
            n_real = len(ids)
            src: list[int] = []
            dst: list[int] = []
            for s, nbrs in self._adj.items():
                si = index.get(s)
                if si is None:
                    si = index[s] = len(ids)
                    ids.append(s)
                for t in nbrs:
                    ti = index.get(t)
                    if ti is None:
                        ti = index[t] = len(ids)
                        ids.append(t)
                    src.append(si)
                    dst.append(ti)
            node_type = np.zeros(len(ids), dtype=np.uint8)
            node_type[n_real:] = ENTITY_CODE_GHOST
            ne = len(src)
            c = csrmod.from_arrays(ids, node_type, np.asarray(src, dtype=np.int32), np.asarray(dst, dtype=np.int32), np.zeros(ne, dtype=np.uint8),
                                   np.full(ne, csrmod.EDGE_TRAVERSABLE, dtype=np.uint8), n_real=n_real)
            c.index = index
            self._csr = c
        return self._csr

    @property
    def device_graph(self):
        if self._dg is None:
            from .engine import DeviceGraph

            self._dg = DeviceGraph.upload(self.csr, self._device)
        return self._dg

    # ── traversals (GPU) ────────────────────────────────────────────────
    def bfs(self, source: str, max_depth: int = 4) -> list[list[str]]:
        """One path per node reached within ``max_depth`` hops, in discovery order (``:67-83``)."""
        if source not in self._nodes:
            return []
        c = self.csr
        res = self.device_graph.bfs_many(np.asarray([c.index[source]], dtype=np.int32), max_depth, False)
        ids = c.node_ids
        nodes, parent = res.slice(0).tolist(), res.aux(0, "parent").tolist()
        paths: list[list[str]] = []
        for i, u in enumerate(nodes):
            p = parent[i]                      # queue position of the first discoverer: 0 = the source, k > 0 = emitted entry k-1
            paths.append(([source] if p <= 0 else paths[p - 1]) + [ids[u]])
        return paths

    def shortest_path(self, source: str, target: str) -> list[str] | None:
        if source not in self._nodes or target not in self._nodes:
            return None
        if source == target:
            return [source]
        c = self.csr
        s, t = c.index[source], c.index[target]
        res = self.device_graph.shortest_path_many(np.asarray([s], dtype=np.int32), np.asarray([t], dtype=np.int32))
        if not (int(res.flags[0]) & _lib.QFLAG_TARGET_FOUND):
            return None
        nodes, parent = res.slice(0), res.aux(0, "parent")
        i = len(nodes) - 1
        while nodes[i] != t:
            i -= 1
        path = []
        while i >= 0:
            path.append(c.node_ids[int(nodes[i])])
            i = int(parent[i])
        return path[::-1]

    def bottleneck_nodes(self, top_n: int = 5) -> list[tuple[str, float]]:
        """Sampled betweenness approximation (``:127-155``): BFS from the first 50 nodes, +1 for every node strictly inside a
        first-discoverer path; normalised by the grand total, highest first, ties in node order."""
        if not self._nodes:
            return []
        c = self.csr
        sample = np.arange(min(_SAMPLE, c.n_real), dtype=np.int32)
        return rank_bottlenecks(c, self.device_graph.bottleneck_scores(sample), top_n)

    @classmethod
    def from_unified_graph(cls, graph, device: int = 0) -> "B200Backend":
        return from_unified_graph(graph, device=device)


def rank_bottlenecks(c: csrmod.HostCSR, scores: np.ndarray, top_n: int) -> list[tuple[str, float]]:
    """Normalise and rank raw device scores the way the reference does; a scored id without a node record is the
    reference's ``KeyError`` (its score table only holds recorded nodes)."""
    ghosts = np.flatnonzero(scores[c.n_real:])
    if ghosts.size:
        raise KeyError(c.node_ids[c.n_real + int(ghosts[0])])
    real = scores[: c.n_real].astype(np.float64)
    total = float(real.sum()) or 1.0
    order = np.argsort(-real, kind="stable")[:top_n]
    return [(c.node_ids[int(i)], float(real[int(i)]) / total) for i in order]


def get_backend(backend: str = "b200", **kwargs) -> B200Backend:
    """``get_backend("b200")`` — the name a ``--graph-backend b200`` flag would pass (reference cli/options_surfaces.py:161-162;
    factory ``graph_backend.py:245-259``).  There is no CPU fallback here: other names are an error."""
    if backend in ("b200", "auto"):
        return B200Backend(**kwargs)
    raise ValueError(f"unknown graph backend {backend!r}: this package only provides 'b200' (no CPU fallback)")


def from_context_graph(context_graph_data: dict, backend: str = "b200", **kwargs) -> B200Backend:
    """Serialised context graph → backend (``:262-278``): undirected edges, kind / label / weight kept."""
    graph = get_backend(backend, **kwargs)
    for node in context_graph_data.get("nodes", []):
        graph.add_node(node_id=node["id"], kind=node.get("kind", ""), label=node.get("label", ""))
    for edge in context_graph_data.get("edges", []):
        graph.add_edge(source=edge["source"], target=edge["target"], kind=edge.get("kind", ""), weight=edge.get("weight", 1.0))
    return graph


def from_unified_graph(ug, backend: str = "b200", **kwargs) -> B200Backend:
    """UnifiedGraph → backend (``:281-306``): directed edges one-way, bidirectional edges both ways."""
    graph = get_backend(backend, **kwargs)
    for node in ug.nodes.values():
        graph.add_node(node_id=node.id, kind=enum_value(node.entity_type), label=node.label, severity=node.severity, risk_score=node.risk_score)
    for edge in ug.edges:
        bidirectional = getattr(edge, "is_bidirectional", None)
        if bidirectional is None:
            bidirectional = getattr(edge, "direction", "directed") == "bidirectional"
        graph.add_edge(source=edge.source, target=edge.target, kind=enum_value(edge.relationship), weight=getattr(edge, "weight", 1.0), directed=not bidirectional)
    return graph

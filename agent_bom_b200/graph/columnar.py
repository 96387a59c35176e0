"""Columnar graph build: inventory report → index arrays → CSR, with no record object per node or edge.

The reference builder allocates one ``UnifiedNode`` and one ``UnifiedEdge`` per element (provenance, timestamps, tags …;
``/root/reference/src/agent_bom/graph/builder.py:41-898``, ≈6 ms per agent), which caps the estate size long before the
GPU does.  ``ColumnSink`` receives the same ``node(...)`` / ``edge(...)`` stream the light builder emits
(``graph/builder.py``) and keeps COLUMNS: id strings + one ``id → index`` dict, entity codes, labels, severities, risk
scores, a sparse attribute map, and the edge stream as four integer arrays with first-wins de-duplication on
``(source, target, relationship)`` — the merge rules of ``UnifiedGraph.add_node`` / ``add_edge`` (reference
``graph/container.py:116-198``) applied to columns.  ``ColumnarGraph`` is a ``UnifiedGraph`` whose CSR is built straight
from those arrays; ``graph.nodes`` / ``graph.edges`` are lazy views that synthesise a ``UnifiedNode`` / ``UnifiedEdge`` only
for the element a caller touches (the ≤ 500 nodes of a ``traverse_subgraph`` answer, the hops of an attack path).
Identity with the object path — and through it with the reference builder — is pinned by ``tests/test_builder_identity.py``.
"""

from __future__ import annotations

from array import array
from collections.abc import Mapping, Sequence
from typing import Any

import numpy as np

from . import csr as csrmod
from .container import UnifiedGraph
from .csr import EDGE_BIDIRECTIONAL, EDGE_TRAVERSABLE, ENTITY_CODE_OTHER
from .model import UnifiedEdge, UnifiedNode
from .schema import ENTITY_CODE, ENTITY_VALUES, REL_CODE, REL_CODE_OTHER, RELATIONSHIP_VALUES, SEVERITY_RANK, EntityType, RelationshipType, enum_value


_REL_OF: dict = {}          # relationship (enum member or string) → 5-bit code, memoised


class ColumnSink:
    def __init__(self, *, scan_id: str = "", tenant_id: str = "", device: int = 0):
        self.scan_id, self.tenant_id, self.device = scan_id, tenant_id, device
        self.ids: list[str] = []
        self.index: dict[str, int] = {}
        self.types = bytearray()
        self.labels: list[str] = []
        self.severity: list[str] = []
        self.risk = array("d")
        self.attrs: dict[int, dict[str, Any]] = {}          # sparse: only nodes that carry attributes
        self.sources: list[str] = []                        # first data-source tag per node
        self.more_sources: dict[int, list[str]] = {}        # sparse: further tags (unioned on merge)
        self.ghosts: dict[str, int] = {}                    # edge endpoints without a node record, first-seen order
        self.e_src, self.e_dst = array("q"), array("q")     # provisional: ghosts are -(k+1) until finish()
        self.e_rel, self.e_flags = bytearray(), bytearray()
        self.e_weight = array("d")
        self._seen: set[tuple[int, int, int]] = set()

    # ── UnifiedGraph.add_node on columns (container.py:116-144) ──
    def node(self, nid: str, et, label: str, source: str, *, severity: str = "", risk_score: float = 0.0, attributes: dict[str, Any] | None = None) -> None:
        i = self.index.get(nid)
        if i is None:
            i = self.index[nid] = len(self.ids)
            self.ids.append(nid)
            self.types.append(ENTITY_CODE.get(enum_value(et), ENTITY_CODE_OTHER))
            self.labels.append(label)
            self.severity.append(severity)
            self.risk.append(float(risk_score or 0.0))
            self.sources.append(source)
            if attributes:
                self.attrs[i] = attributes
            return
        if attributes:
            have = self.attrs.get(i)
            if have is None:
                self.attrs[i] = dict(attributes)
            else:
                have.update(attributes)
        if SEVERITY_RANK.get(severity, 0) > SEVERITY_RANK.get(self.severity[i], 0):
            self.severity[i] = severity
        if (risk_score or 0.0) > self.risk[i]:
            self.risk[i] = float(risk_score)
        if source != self.sources[i]:
            extra = self.more_sources.setdefault(i, [])
            if source not in extra:
                extra.append(source)

    def _endpoint(self, nid: str) -> int:
        i = self.index.get(nid)
        if i is not None:
            return i
        k = self.ghosts.get(nid)
        if k is None:
            k = self.ghosts[nid] = len(self.ghosts)
        return -(k + 1)

    # ── UnifiedGraph.add_edge on columns: first (source, target, relationship) wins (container.py:146-172) ──
    def edge(self, src: str, dst: str, rel, *, direction: str = "directed", weight: float = 1.0, traversable: bool = True) -> None:
        index = self.index
        s = index.get(src)
        if s is None:
            s = self._endpoint(src)
        d = index.get(dst)
        if d is None:
            d = self._endpoint(dst)
        r = _REL_OF.get(rel)
        if r is None:
            r = _REL_OF[rel] = REL_CODE.get(enum_value(rel), REL_CODE_OTHER)
        key = (s, d, r)
        if key in self._seen:
            return
        self._seen.add(key)
        self.e_src.append(s); self.e_dst.append(d)
        self.e_rel.append(r)
        self.e_flags.append((EDGE_TRAVERSABLE if traversable else 0) | (EDGE_BIDIRECTIONAL if direction == "bidirectional" else 0))
        self.e_weight.append(weight)

    def has_node(self, nid: str) -> bool:
        return nid in self.index

    def finish(self) -> "ColumnarGraph":
        self._seen = set()
        return ColumnarGraph(self)


class _NodeView(Mapping):
    """``graph.nodes``: id → record, records synthesised on access (and cached, so attribute edits stick)."""

    def __init__(self, cols: ColumnSink):
        self._c = cols
        self._cache: dict[int, UnifiedNode] = {}

    def _record(self, i: int) -> UnifiedNode:
        rec = self._cache.get(i)
        if rec is None:
            c = self._c
            code = c.types[i]
            et = ENTITY_VALUES[code] if code < len(ENTITY_VALUES) else "other"
            try:
                et = EntityType(et)
            except ValueError:
                pass
            rec = self._cache[i] = UnifiedNode(id=c.ids[i], entity_type=et, label=c.labels[i], risk_score=c.risk[i], severity=c.severity[i],
                                               attributes=c.attrs.setdefault(i, {}), data_sources=[c.sources[i], *c.more_sources.get(i, ())])
        return rec

    def __getitem__(self, nid: str) -> UnifiedNode:
        return self._record(self._c.index[nid])

    def get(self, nid, default=None):
        i = self._c.index.get(nid)
        return default if i is None else self._record(i)

    def __contains__(self, nid) -> bool:
        return nid in self._c.index

    def __iter__(self):
        return iter(self._c.ids)

    def __len__(self) -> int:
        return len(self._c.ids)

    def values(self):
        return (self._record(i) for i in range(len(self._c.ids)))

    def items(self):
        return ((self._c.ids[i], self._record(i)) for i in range(len(self._c.ids)))


class _EdgeView(Sequence):
    """``graph.edges``: the edge stream as records, synthesised per access from the index arrays."""

    def __init__(self, graph: "ColumnarGraph"):
        self._g = graph

    def __len__(self) -> int:
        return int(self._g._e_src.shape[0])

    def _record(self, i: int) -> UnifiedEdge:
        g = self._g
        ids = g._all_ids
        code = int(g._e_rel[i])
        rel = RELATIONSHIP_VALUES[code] if code < len(RELATIONSHIP_VALUES) else "other"
        try:
            rel = RelationshipType(rel)
        except ValueError:
            pass
        fl = int(g._e_flags[i])
        return UnifiedEdge(source=ids[int(g._e_src[i])], target=ids[int(g._e_dst[i])], relationship=rel,
                           direction="bidirectional" if fl & EDGE_BIDIRECTIONAL else "directed", weight=float(g._e_weight[i]), traversable=bool(fl & EDGE_TRAVERSABLE))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._record(k) for k in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return self._record(i)

    def __iter__(self):
        return (self._record(i) for i in range(len(self)))


class ColumnarGraph(UnifiedGraph):
    """A ``UnifiedGraph`` over columns: the CSR is ready at construction, records exist only for what is touched.

    Mutating it (``add_node`` / ``add_edge``) first materialises ordinary records, after which it behaves like any ``UnifiedGraph``."""

    def __init__(self, cols: ColumnSink):
        super().__init__(scan_id=cols.scan_id, tenant_id=cols.tenant_id, device=cols.device)
        n_real, n_ghost = len(cols.ids), len(cols.ghosts)
        src = np.frombuffer(cols.e_src, dtype=np.int64) if len(cols.e_src) else np.zeros(0, np.int64)
        dst = np.frombuffer(cols.e_dst, dtype=np.int64) if len(cols.e_dst) else np.zeros(0, np.int64)
        self._e_src = np.where(src < 0, n_real - 1 - src, src).astype(np.int32)       # ghost k → n_real + k
        self._e_dst = np.where(dst < 0, n_real - 1 - dst, dst).astype(np.int32)
        self._e_rel = np.frombuffer(bytes(cols.e_rel), dtype=np.uint8)
        self._e_flags = np.frombuffer(bytes(cols.e_flags), dtype=np.uint8)
        self._e_weight = np.frombuffer(cols.e_weight, dtype=np.float64) if len(cols.e_weight) else np.zeros(0, np.float64)
        self._all_ids = cols.ids + list(cols.ghosts) if n_ghost else cols.ids
        types = np.concatenate([np.frombuffer(bytes(cols.types), dtype=np.uint8), np.full(n_ghost, 255, dtype=np.uint8)]) if n_ghost else np.frombuffer(bytes(cols.types), dtype=np.uint8)
        self._cols = cols
        self.nodes = _NodeView(cols)
        self.edges = _EdgeView(self)
        self._csr = csrmod.from_arrays(self._all_ids, types, self._e_src, self._e_dst, self._e_rel, self._e_flags, n_real=n_real)

    def _materialise(self) -> None:
        if isinstance(self.nodes, _NodeView):
            nodes, edges = dict(self.nodes.items()), list(self.edges)
            self.nodes, self.edges = nodes, edges
            for i, e in enumerate(edges):
                key = (e.source, e.target, enum_value(e.relationship))
                self._edge_keys.add(key)
                self._edge_index.setdefault(key, i)

    def add_node(self, node) -> None:
        self._materialise()
        super().add_node(node)

    def add_edge(self, edge) -> None:
        self._materialise()
        super().add_edge(edge)

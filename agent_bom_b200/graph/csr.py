"""Host-side CSR export: typed graph -> dense index space -> adjacency / reverse-adjacency CSR.

The inventory graph reaches the device as two CSRs whose rows reproduce
``graph.adjacency[u]`` and ``graph.reverse_adjacency[u]`` *in list order*
(reference ``/root/reference/src/agent_bom/graph/container.py:146-198``); that
order is what makes the engine's BFS discovery order identical to the
reference's.  Node index = insertion order of ``graph.nodes``; edge endpoints
without a node record ("ghosts", e.g. runtime edges — reference
``graph/builder.py:804-819``) get indices after the real nodes.

The stable counting sort itself runs in the C++ library
(``abb_csr_build_host``); this module only maps strings to indices.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .. import _lib
from .schema import ENTITY_CODE, ENTITY_CODE_GHOST, REL_CODE, REL_CODE_OTHER, enum_value

EDGE_TRAVERSABLE, EDGE_BIDIRECTIONAL = 1, 2
ENTITY_CODE_OTHER = 254


@dataclass
class HostCSR:
    """Index-space view of one graph: string table + edge stream + both CSRs (numpy, host memory)."""

    node_ids: list[str]
    n_real: int
    node_type: np.ndarray          # uint8 [n]
    node_rank: np.ndarray          # int32 [n] rank of the id string among all ids
    src: np.ndarray                # int32 [E] graph.edges order
    dst: np.ndarray
    rel: np.ndarray                # uint8
    flags: np.ndarray              # uint8 EDGE_*
    fwd_off: np.ndarray = field(default=None)
    fwd_nbr: np.ndarray = field(default=None)
    fwd_meta: np.ndarray = field(default=None)
    fwd_eid: np.ndarray = field(default=None)
    rev_off: np.ndarray = field(default=None)
    rev_nbr: np.ndarray = field(default=None)
    rev_meta: np.ndarray = field(default=None)
    rev_eid: np.ndarray = field(default=None)
    index: dict[str, int] | None = None

    @property
    def n_nodes(self) -> int:
        return int(self.node_type.shape[0])

    @property
    def n_edges(self) -> int:
        return int(self.src.shape[0])

    @property
    def n_entries(self) -> int:
        return int(self.fwd_nbr.shape[0])

    def idx(self, node_id: str) -> int:
        if self.index is None:
            self.index = {nid: i for i, nid in enumerate(self.node_ids)}
        return self.index.get(node_id, -1)

    def nbytes(self) -> int:
        return sum(getattr(self, n).nbytes for n in ("fwd_off", "fwd_nbr", "fwd_meta", "fwd_eid", "rev_off", "rev_nbr", "rev_meta", "rev_eid", "node_type", "node_rank"))

    def c_struct(self) -> _lib.Csr:
        c = _lib.Csr()
        c.n_nodes = self.n_nodes
        c.n_entries = self.n_entries
        for name in ("fwd_off", "fwd_nbr", "fwd_meta", "fwd_eid", "rev_off", "rev_nbr", "rev_meta", "rev_eid", "node_type", "node_rank"):
            arr = getattr(self, name)
            assert arr.flags["C_CONTIGUOUS"]
            setattr(c, name, arr.ctypes.data)
        return c


def string_rank(ids: list[str]) -> np.ndarray:
    """rank[i] = position of ids[i] in sorted(ids) (the order of the reference's ``sorted()`` over id strings)."""
    order = sorted(range(len(ids)), key=ids.__getitem__)
    rank = np.empty(len(ids), dtype=np.int32)
    rank[np.asarray(order, dtype=np.int64)] = np.arange(len(ids), dtype=np.int32)
    return rank


def build_rows(n_nodes: int, src, dst, rel, flags):
    """Run the library's stable counting sort; returns the eight CSR arrays."""
    lib = _lib.load()
    src = np.ascontiguousarray(src, dtype=np.int32)
    dst = np.ascontiguousarray(dst, dtype=np.int32)
    rel = np.ascontiguousarray(rel, dtype=np.uint8)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    ne = int(src.shape[0])
    m = int(lib.abb_csr_entries(ne, flags.ctypes.data)) if ne else 0
    out = {
        "fwd_off": np.zeros(n_nodes + 1, dtype=np.uint32), "fwd_nbr": np.zeros(m, dtype=np.int32), "fwd_meta": np.zeros(m, dtype=np.uint8),
        "fwd_eid": np.zeros(m, dtype=np.uint32), "rev_off": np.zeros(n_nodes + 1, dtype=np.uint32), "rev_nbr": np.zeros(m, dtype=np.int32),
        "rev_meta": np.zeros(m, dtype=np.uint8), "rev_eid": np.zeros(m, dtype=np.uint32),
    }
    _lib.check(lib.abb_csr_build_host(
        n_nodes, ne, src.ctypes.data, dst.ctypes.data, rel.ctypes.data, flags.ctypes.data,
        out["fwd_off"].ctypes.data, out["fwd_nbr"].ctypes.data, out["fwd_meta"].ctypes.data, out["fwd_eid"].ctypes.data,
        out["rev_off"].ctypes.data, out["rev_nbr"].ctypes.data, out["rev_meta"].ctypes.data, out["rev_eid"].ctypes.data,
    ))
    return out


def from_arrays(node_ids, node_type, src, dst, rel, flags, *, n_real: int | None = None, node_rank=None) -> HostCSR:
    """CSR from an index-space edge stream (used by the estate generator and the golden fixtures)."""
    node_type = np.ascontiguousarray(node_type, dtype=np.uint8)
    n = int(node_type.shape[0])
    ids = list(node_ids) if node_ids is not None else None
    if node_rank is None:
        node_rank = string_rank(ids) if ids is not None else np.arange(n, dtype=np.int32)
    rows = build_rows(n, src, dst, rel, flags)
    return HostCSR(
        node_ids=ids if ids is not None else [], n_real=n if n_real is None else n_real, node_type=node_type,
        node_rank=np.ascontiguousarray(node_rank, dtype=np.int32),
        src=np.ascontiguousarray(src, dtype=np.int32), dst=np.ascontiguousarray(dst, dtype=np.int32),
        rel=np.ascontiguousarray(rel, dtype=np.uint8), flags=np.ascontiguousarray(flags, dtype=np.uint8), **rows,
    )


def from_unified_graph(graph) -> HostCSR:
    """Export any UnifiedGraph-shaped object (ours or the reference's: ``.nodes`` dict, ``.edges`` list)."""
    ids = list(graph.nodes.keys())
    index = {nid: i for i, nid in enumerate(ids)}
    n_real = len(ids)
    types = [ENTITY_CODE.get(enum_value(n.entity_type), ENTITY_CODE_OTHER) for n in graph.nodes.values()]
    ne = len(graph.edges)
    src = np.empty(ne, dtype=np.int32)
    dst = np.empty(ne, dtype=np.int32)
    rel = np.empty(ne, dtype=np.uint8)
    flags = np.empty(ne, dtype=np.uint8)
    for i, e in enumerate(graph.edges):
        s = index.get(e.source)
        if s is None:
            s = index[e.source] = len(ids)
            ids.append(e.source)
        t = index.get(e.target)
        if t is None:
            t = index[e.target] = len(ids)
            ids.append(e.target)
        src[i], dst[i] = s, t
        rel[i] = REL_CODE.get(enum_value(e.relationship), REL_CODE_OTHER)
        flags[i] = (EDGE_TRAVERSABLE if e.traversable else 0) | (EDGE_BIDIRECTIONAL if e.direction == "bidirectional" else 0)
    node_type = np.asarray(types + [ENTITY_CODE_GHOST] * (len(ids) - n_real), dtype=np.uint8)
    csr = from_arrays(ids, node_type, src, dst, rel, flags, n_real=n_real)
    csr.index = index
    return csr

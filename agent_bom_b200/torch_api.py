"""Device-buffer form of the engine for callers that keep their data in HBM (torch tensors).

torch is plumbing here: it owns device memory, streams and (in ``dist.py``) the
NCCL process group; every kernel is the library's own (``abb_walk_launch``,
``abb_paths_*_launch``).  All launches are asynchronous on the given / current
torch stream.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .engine import DeviceGraph
from .graph.schema import N_ENTITY_TYPES


def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(stream: torch.cuda.Stream | None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def frontier_signatures(dg: DeviceGraph, spec, roots: torch.Tensor, stream: torch.cuda.Stream | None = None) -> torch.Tensor:
    """48-bit signature of every source's depth-1 frontier (bit 47 set = walked individually, the key is then the query index); equal signatures share a traversal."""
    assert roots.dtype == torch.int32 and roots.is_cuda
    sig = torch.empty(roots.shape[0], dtype=torch.int64, device=roots.device)
    _lib.check(_lib.load().abb_walk_signatures(dg.handle, C.byref(spec), _ptr(roots), int(roots.shape[0]), _ptr(sig), _stream_ptr(stream)))
    return sig


def shard_by_signature(sig: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    """Boolean mask of the sources this rank owns: all sources of one frontier group land on the same rank."""
    return ((sig & 0x7FFFFFFFFFFFFFFF) % world) == rank


class DeviceWalk:
    """Reusable output buffers + launch for one walk spec over batches of up to ``max_queries`` queries."""

    def __init__(self, dg: DeviceGraph, spec, max_queries: int, node_cap: int, edge_cap: int = 0):
        self.dg, self.spec = dg, spec
        dev = torch.device("cuda", dg.device)
        fl = spec.flags
        q = max(int(max_queries), 1)
        self.max_queries = q
        self.q_start = torch.empty(q, dtype=torch.int64, device=dev)
        self.q_count = torch.empty(q, dtype=torch.int32, device=dev)
        self.q_maxd = torch.empty(q, dtype=torch.int32, device=dev)
        self.q_flags = torch.empty(q, dtype=torch.int32, device=dev)
        self.q_hist = torch.empty(q * N_ENTITY_TYPES, dtype=torch.int32, device=dev) if fl & _lib.WALK_HIST else None
        self.q_estart = torch.empty(q, dtype=torch.int64, device=dev) if fl & _lib.WALK_EDGES else None
        self.q_ecount = torch.empty(q, dtype=torch.int64, device=dev) if fl & _lib.WALK_EDGES else None
        self.totals = torch.zeros(2, dtype=torch.int64, device=dev)
        self.node_cap = self.edge_cap = 0
        self.nodes = self.parent = self.depth = self.edges = None
        self.reserve(node_cap, edge_cap)

    def reserve(self, node_cap: int, edge_cap: int = 0) -> None:
        dev = self.q_start.device
        fl = self.spec.flags
        if node_cap > self.node_cap:
            self.node_cap = int(node_cap)
            self.nodes = torch.empty(self.node_cap, dtype=torch.int32, device=dev)
            self.parent = torch.empty(self.node_cap, dtype=torch.int32, device=dev) if fl & _lib.WALK_PARENTS else None
            self.depth = torch.empty(self.node_cap, dtype=torch.int32, device=dev) if fl & _lib.WALK_DEPTHS else None
        if (fl & _lib.WALK_EDGES) and edge_cap > self.edge_cap:
            self.edge_cap = int(edge_cap)
            self.edges = torch.empty(self.edge_cap, dtype=torch.int32, device=dev)

    def launch(self, roots: torch.Tensor, n_queries: int | None = None, root_off: torch.Tensor | None = None, targets: torch.Tensor | None = None,
               stream: torch.cuda.Stream | None = None) -> None:
        nq = int(n_queries if n_queries is not None else (root_off.shape[0] - 1 if root_off is not None else roots.shape[0]))
        assert nq <= self.max_queries and roots.dtype == torch.int32 and roots.is_cuda
        io = _lib.WalkIO()
        io.n_queries = nq
        io.roots, io.root_off, io.targets = _ptr(roots), _ptr(root_off), _ptr(targets)
        io.q_start, io.q_count, io.q_maxd, io.q_flags = _ptr(self.q_start), _ptr(self.q_count), _ptr(self.q_maxd), _ptr(self.q_flags)
        io.q_estart, io.q_ecount, io.q_hist = _ptr(self.q_estart), _ptr(self.q_ecount), _ptr(self.q_hist)
        io.nodes, io.parent, io.depth, io.node_cap = _ptr(self.nodes), _ptr(self.parent), _ptr(self.depth), self.node_cap
        io.edges, io.edge_cap, io.totals = _ptr(self.edges), self.edge_cap, _ptr(self.totals)
        _lib.check(_lib.load().abb_walk_launch(self.dg.handle, C.byref(self.spec), C.byref(io), _stream_ptr(stream)))

    def needed(self) -> tuple[int, int]:
        """(node slots, edge slots) the last launch needed — synchronises."""
        t = self.totals.cpu()
        return int(t[0]), int(t[1])

    def launch_fitted(self, roots: torch.Tensor, **kw) -> tuple[int, int]:
        """Launch, and if the arenas were too small grow them to the exact need and launch again."""
        self.launch(roots, **kw)
        need = self.needed()
        if need[0] > self.node_cap or need[1] > self.edge_cap:
            self.reserve(need[0], need[1])
            self.launch(roots, **kw)
            need = self.needed()
        return need


class DevicePaths:
    """Exposure-path rows for batches of findings resident on the device (count pass + fill pass)."""

    def __init__(self, dg: DeviceGraph, max_findings: int, row_cap: int = 0):
        self.dg = dg
        dev = torch.device("cuda", dg.device)
        self.f_off = torch.zeros(max(int(max_findings), 1) + 1, dtype=torch.int64, device=dev)
        self.max_findings = max(int(max_findings), 1)
        self.row_cap = 0
        self.hops = self.rels = self.ncred = self.ntool = None
        self.reserve(max(row_cap, 1))

    def reserve(self, row_cap: int) -> None:
        if row_cap > self.row_cap:
            dev = self.f_off.device
            self.row_cap = int(row_cap)
            self.hops = torch.empty(self.row_cap * 4, dtype=torch.int32, device=dev)
            self.rels = torch.empty(self.row_cap * 4, dtype=torch.int8, device=dev)   # 4 bytes per row: 3 relationships + pad
            self.ncred = torch.empty(self.row_cap, dtype=torch.int32, device=dev)
            self.ntool = torch.empty(self.row_cap, dtype=torch.int32, device=dev)

    def _io(self, findings: torch.Tensor, n: int) -> _lib.PathsIO:
        io = _lib.PathsIO()
        io.n_findings = n
        io.findings, io.f_off = _ptr(findings), _ptr(self.f_off)
        io.hops, io.rels, io.ncred, io.ntool, io.row_cap = _ptr(self.hops), _ptr(self.rels), _ptr(self.ncred), _ptr(self.ntool), self.row_cap
        return io

    def count(self, findings: torch.Tensor, n: int | None = None, stream=None) -> None:
        n = int(findings.shape[0] if n is None else n)
        assert n <= self.max_findings
        io = self._io(findings, n)
        _lib.check(_lib.load().abb_paths_count_launch(self.dg.handle, C.byref(io), _stream_ptr(stream)))

    def fill(self, findings: torch.Tensor, n: int | None = None, stream=None) -> None:
        n = int(findings.shape[0] if n is None else n)
        io = self._io(findings, n)
        _lib.check(_lib.load().abb_paths_fill_launch(self.dg.handle, C.byref(io), _stream_ptr(stream)))

    def total_rows(self, n: int) -> int:
        return int(self.f_off[n].item())

    def run_fitted(self, findings: torch.Tensor, n: int | None = None) -> int:
        n = int(findings.shape[0] if n is None else n)
        self.count(findings, n)
        rows = self.total_rows(n)
        self.reserve(rows)
        self.fill(findings, n)
        return rows

"""ctypes front-end of the CPU oracle (oracle.c) + numpy restatement of the adjacency model.

TEST INFRASTRUCTURE ONLY — see the header of oracle.c.  Only tests/,
``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs import this module.

``build_csr`` restates ``UnifiedGraph.add_edge`` (reference
``/root/reference/src/agent_bom/graph/container.py:146-198``): adjacency and
reverse-adjacency lists in insertion order, reversed copies for bidirectional
edges.  It is deliberately independent of the product's C++ CSR builder.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"

META_TRAV, META_FIRSTPAIR, META_REVCOPY = 0x20, 0x40, 0x80
FLAG_TRAVERSABLE, FLAG_BIDIRECTIONAL = 1, 2
GHOST = 255
NHIST = 24


def build_lib(force: bool = False) -> Path:
    src = HERE / "oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-D_GNU_SOURCE", "-o", str(LIB_PATH), str(src)]
        try:
            subprocess.run(cmd, check=True, capture_output=True, text=True)
        except subprocess.CalledProcessError:
            cmd.remove("-march=native")
            subprocess.run(cmd, check=True, capture_output=True, text=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_lib()
        L = C.CDLL(str(LIB_PATH))
        vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
        L.orc_impact_many.restype = vp
        L.orc_impact_many.argtypes = [vp, vp, i64, i32, C.c_int]
        L.orc_bfs_many.restype = vp
        L.orc_bfs_many.argtypes = [vp, vp, i64, i32, i32, C.c_int]
        L.orc_reachable_many.restype = vp
        L.orc_reachable_many.argtypes = [vp, vp, i64, i32, i32, C.c_int]
        L.orc_distances_many.restype = vp
        L.orc_distances_many.argtypes = [vp, vp, i64, u32, C.c_int]
        L.orc_traverse_many.restype = vp
        L.orc_traverse_many.argtypes = [vp, vp, vp, i64, i32, i32, i64, i64, u32, i32, i32, C.c_int]
        L.orc_shortest_path.restype = i64
        L.orc_shortest_path.argtypes = [vp, i32, i32, vp, i64]
        L.orc_derived_paths.restype = vp
        L.orc_derived_paths.argtypes = [vp, vp, i64, vp, C.c_int]
        for name in ("total", "total_edges"):
            f = getattr(L, f"orc_result_{name}")
            f.restype, f.argtypes = i64, [vp]
        for name in ("off", "nodes", "aux", "eoff", "edges", "hist", "maxd", "flags", "edge_count"):
            f = getattr(L, f"orc_result_{name}")
            f.restype, f.argtypes = vp, [vp]
        L.orc_result_counters.argtypes = [vp, vp]
        L.orc_result_free.argtypes = [vp]
        L.orc_paths_count.restype, L.orc_paths_count.argtypes = i64, [vp]
        for name in ("hops", "rels", "ncred", "ntool"):
            f = getattr(L, f"orc_paths_{name}")
            f.restype, f.argtypes = vp, [vp]
        L.orc_paths_free.argtypes = [vp]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


class _CGraph(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_int32),
        ("n_entries", C.c_int64),
        ("fwd_off", C.c_void_p), ("fwd_nbr", C.c_void_p), ("fwd_meta", C.c_void_p), ("fwd_eid", C.c_void_p),
        ("rev_off", C.c_void_p), ("rev_nbr", C.c_void_p), ("rev_meta", C.c_void_p), ("rev_eid", C.c_void_p),
        ("node_type", C.c_void_p),
    ]


def _stable_rows(row: np.ndarray, order_key: np.ndarray, n_nodes: int):
    """Stable grouping of adjacency entries by row, preserving insertion order."""
    perm = np.lexsort((order_key, row))
    counts = np.bincount(row, minlength=n_nodes).astype(np.int64)
    off = np.zeros(n_nodes + 1, dtype=np.uint32)
    off[1:] = np.cumsum(counts)
    return perm, off


@dataclass
class OracleGraph:
    """CSR pair + node types, kept alive for the C side."""

    n_nodes: int
    fwd_off: np.ndarray
    fwd_nbr: np.ndarray
    fwd_meta: np.ndarray
    fwd_eid: np.ndarray
    rev_off: np.ndarray
    rev_nbr: np.ndarray
    rev_meta: np.ndarray
    rev_eid: np.ndarray
    node_type: np.ndarray
    _c: _CGraph | None = None

    @property
    def n_entries(self) -> int:
        return int(self.fwd_nbr.shape[0])

    def cptr(self):
        if self._c is None:
            g = _CGraph()
            g.n_nodes = self.n_nodes
            g.n_entries = self.n_entries
            for name in ("fwd_off", "fwd_nbr", "fwd_meta", "fwd_eid", "rev_off", "rev_nbr", "rev_meta", "rev_eid", "node_type"):
                arr = getattr(self, name)
                assert arr.flags["C_CONTIGUOUS"]
                setattr(g, name, arr.ctypes.data)
            self._c = g
        return C.addressof(self._c)


def build_csr(n_nodes: int, src, dst, rel, flags, node_type) -> OracleGraph:
    """Edge stream (graph.edges order) -> adjacency / reverse-adjacency CSR (container.py:146-198)."""
    src = np.ascontiguousarray(src, dtype=np.int64)
    dst = np.ascontiguousarray(dst, dtype=np.int64)
    rel = np.ascontiguousarray(rel, dtype=np.uint8)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    ne = src.shape[0]
    idx = np.arange(ne, dtype=np.int64)
    bid = (flags & FLAG_BIDIRECTIONAL) != 0
    base_meta = (rel & 0x1F) | np.where(flags & FLAG_TRAVERSABLE, META_TRAV, 0).astype(np.uint8)
    bidx = idx[bid]
    # forward: original under source, reversed copy under target (right after the original in insertion time)
    f_row = np.concatenate([src, dst[bid]])
    f_nbr = np.concatenate([dst, src[bid]])
    f_key = np.concatenate([2 * idx, 2 * bidx + 1])
    f_meta = np.concatenate([base_meta, base_meta[bid] | META_REVCOPY]).astype(np.uint8)
    # reverse: original under target, reversed copy under source
    r_row = np.concatenate([dst, src[bid]])
    r_nbr = np.concatenate([src, dst[bid]])
    r_key = f_key
    r_meta = f_meta
    fp, f_off = _stable_rows(f_row, f_key, n_nodes)
    rp, r_off = _stable_rows(r_row, r_key, n_nodes)
    nt = np.ascontiguousarray(node_type, dtype=np.uint8)
    assert nt.shape[0] == n_nodes

    def first_pair(off, nbr_sorted, meta_sorted):
        # first entry of each row with a given neighbour (graph.py:492-497 keeps that edge's relationship for the pair)
        row = np.repeat(np.arange(n_nodes, dtype=np.int64), np.diff(off.astype(np.int64)))
        _, first = np.unique(row * n_nodes + nbr_sorted.astype(np.int64), return_index=True)
        out = meta_sorted.copy()
        out[first] |= META_FIRSTPAIR
        return out

    f_meta_sorted = first_pair(f_off, f_nbr[fp], f_meta[fp])
    r_meta_sorted = first_pair(r_off, r_nbr[rp], r_meta[rp])
    return OracleGraph(
        n_nodes=n_nodes,
        fwd_off=f_off, fwd_nbr=np.ascontiguousarray(f_nbr[fp], dtype=np.int32), fwd_meta=np.ascontiguousarray(f_meta_sorted),
        fwd_eid=np.ascontiguousarray(f_key[fp], dtype=np.uint32),
        rev_off=r_off, rev_nbr=np.ascontiguousarray(r_nbr[rp], dtype=np.int32), rev_meta=np.ascontiguousarray(r_meta_sorted),
        rev_eid=np.ascontiguousarray(r_key[rp], dtype=np.uint32),
        node_type=nt,
    )


@dataclass
class WalkResult:
    off: np.ndarray          # int64 [Q+1]
    nodes: np.ndarray        # int32, discovery order
    aux: np.ndarray          # int32 parent position (bfs) or depth
    maxd: np.ndarray
    flags: np.ndarray
    hist: np.ndarray | None = None
    eoff: np.ndarray | None = None
    edges: np.ndarray | None = None
    edge_count: np.ndarray | None = None
    n_exp: int = 0
    m_scan: int = 0
    n_disc: int = 0

    @property
    def algorithmic_bytes(self) -> int:
        """B = 8·N_exp + 6·M_scan + 8·N_disc."""
        return 8 * self.n_exp + 6 * self.m_scan + 8 * self.n_disc

    def slice(self, q: int):
        a, b = int(self.off[q]), int(self.off[q + 1])
        return self.nodes[a:b], self.aux[a:b]


def _np_from(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def _collect(res, nq: int, hist=False, edges=False) -> WalkResult:
    L = lib()
    total = L.orc_result_total(res)
    out = WalkResult(
        off=_np_from(L.orc_result_off(res), nq + 1, np.int64),
        nodes=_np_from(L.orc_result_nodes(res), total, np.int32),
        aux=_np_from(L.orc_result_aux(res), total, np.int32),
        maxd=_np_from(L.orc_result_maxd(res), nq, np.int32),
        flags=_np_from(L.orc_result_flags(res), nq, np.int32),
    )
    if hist:
        out.hist = _np_from(L.orc_result_hist(res), nq * NHIST, np.uint32).reshape(nq, NHIST)
    if edges:
        te = L.orc_result_total_edges(res)
        out.eoff = _np_from(L.orc_result_eoff(res), nq + 1, np.int64)
        out.edges = _np_from(L.orc_result_edges(res), te, np.uint32)
        out.edge_count = _np_from(L.orc_result_edge_count(res), nq, np.int64)
    c3 = (C.c_int64 * 3)()
    L.orc_result_counters(res, c3)
    out.n_exp, out.m_scan, out.n_disc = int(c3[0]), int(c3[1]), int(c3[2])
    L.orc_result_free(res)
    return out


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def impact_many(g: OracleGraph, sources, max_depth: int = 4, threads: int = 0) -> WalkResult:
    s = _i32(sources)
    return _collect(lib().orc_impact_many(g.cptr(), s.ctypes.data, len(s), max_depth, threads), len(s), hist=True)


def bfs_many(g: OracleGraph, sources, max_depth: int = 4, traversable_only: bool = True, threads: int = 0) -> WalkResult:
    s = _i32(sources)
    return _collect(lib().orc_bfs_many(g.cptr(), s.ctypes.data, len(s), max_depth, int(traversable_only), threads), len(s))


def reachable_many(g: OracleGraph, sources, max_depth: int = 6, traversable_only: bool = False, threads: int = 0) -> WalkResult:
    s = _i32(sources)
    return _collect(lib().orc_reachable_many(g.cptr(), s.ctypes.data, len(s), max_depth, int(traversable_only), threads), len(s))


def distances_many(g: OracleGraph, sources, rel_mask: int, threads: int = 0) -> WalkResult:
    s = _i32(sources)
    return _collect(lib().orc_distances_many(g.cptr(), s.ctypes.data, len(s), rel_mask & 0xFFFFFFFF, threads), len(s))


def traverse_many(g: OracleGraph, roots, root_off, *, direction: int, max_depth: int, max_nodes: int = -1, max_edges: int = -1,
                  rel_mask: int = 0xFFFFFFFF, traversable_only: bool = False, include_roots: bool = True, threads: int = 0) -> WalkResult:
    r = _i32(roots)
    ro = np.ascontiguousarray(root_off, dtype=np.int64)
    nq = len(ro) - 1
    res = lib().orc_traverse_many(g.cptr(), r.ctypes.data, ro.ctypes.data, nq, direction, max_depth, max_nodes, max_edges,
                                  rel_mask & 0xFFFFFFFF, int(traversable_only), int(include_roots), threads)
    return _collect(res, nq, edges=True)


def shortest_path(g: OracleGraph, src: int, dst: int):
    cap = g.n_nodes + 1
    buf = np.zeros(cap, dtype=np.int32)
    n = lib().orc_shortest_path(g.cptr(), int(src), int(dst), buf.ctypes.data, cap)
    return None if n == 0 else buf[:n].copy()


@dataclass
class PathRows:
    hops: np.ndarray   # int32 [P,4]  agent, server, vuln_source(-1 when it is the server), finding
    rels: np.ndarray   # int8  [P,3]  rel code per hop pair; -1 pair without edge; -2 not applicable
    ncred: np.ndarray
    ntool: np.ndarray


def derived_paths(g: OracleGraph, findings, node_rank, threads: int = 0) -> PathRows:
    f = _i32(findings)
    rk = _i32(node_rank)
    L = lib()
    p = L.orc_derived_paths(g.cptr(), f.ctypes.data, len(f), rk.ctypes.data, threads)
    n = L.orc_paths_count(p)
    out = PathRows(
        hops=_np_from(L.orc_paths_hops(p), n * 4, np.int32).reshape(n, 4),
        rels=_np_from(L.orc_paths_rels(p), n * 3, np.int8).reshape(n, 3),
        ncred=_np_from(L.orc_paths_ncred(p), n, np.int32),
        ntool=_np_from(L.orc_paths_ntool(p), n, np.int32),
    )
    L.orc_paths_free(p)
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())


def dependency_reach(g: OracleGraph, agents, rel_mask: int, vuln_pkg_mask: int, node_rank, threads: int = 0):
    """compute_dependency_reach (graph/dependency_reach.py:109-166) on top of distances_many.

    Returns (pkg_ids, pkg_off, pkg_agents(sorted by node_rank), pkg_minhop,
             vuln_ids, vuln_poff, vuln_pkgs(sorted by rank), vuln_aoff, vuln_agents(sorted), vuln_minhop).
    """
    ET_PACKAGE, ET_VULN = 2, 8
    agents = _i32(agents)
    node_rank = np.asarray(node_rank)
    w = distances_many(g, agents, rel_mask, threads)
    q_of = np.repeat(np.arange(len(agents)), np.diff(w.off))
    is_pkg = g.node_type[w.nodes] == ET_PACKAGE
    pk, ag, hp = w.nodes[is_pkg], agents[q_of[is_pkg]], w.aux[is_pkg]
    # agents themselves are never packages, so dropping the start node loses nothing
    pkg_ids = np.flatnonzero(g.node_type == ET_PACKAGE).astype(np.int32)
    order = np.lexsort((node_rank[ag], pk))
    pk, ag, hp = pk[order], ag[order], hp[order]
    counts = np.bincount(pk, minlength=g.n_nodes)[pkg_ids]
    pkg_off = np.zeros(len(pkg_ids) + 1, dtype=np.int64)
    pkg_off[1:] = np.cumsum(counts)
    minhop_all = np.zeros(g.n_nodes, dtype=np.int32)
    if len(pk):
        big = np.full(g.n_nodes, np.iinfo(np.int32).max, dtype=np.int64)
        np.minimum.at(big, pk, hp)
        reached = big != np.iinfo(np.int32).max
        minhop_all[reached] = big[reached]
    pkg_minhop = minhop_all[pkg_ids]
    # vulnerabilities: packages attached through affects / vulnerable_to in either list (:201-220)
    vuln_ids = np.flatnonzero(g.node_type == ET_VULN).astype(np.int32)
    pkg_slot = np.full(g.n_nodes, -1, dtype=np.int64)
    pkg_slot[pkg_ids] = np.arange(len(pkg_ids))
    vp_off = [0]; vp = []; va_off = [0]; va = []; vmin = []
    for v in vuln_ids:
        att = set()
        for off, nbr, meta in ((g.fwd_off, g.fwd_nbr, g.fwd_meta), (g.rev_off, g.rev_nbr, g.rev_meta)):
            a, b = int(off[v]), int(off[v + 1])
            for p in range(a, b):
                if (vuln_pkg_mask >> (int(meta[p]) & 0x1F)) & 1 and g.node_type[nbr[p]] == ET_PACKAGE:
                    att.add(int(nbr[p]))
        att_sorted = sorted(att, key=lambda u: node_rank[u])
        vp.extend(att_sorted); vp_off.append(len(vp))
        agents_u = set(); best = None
        for pkg in att_sorted:
            sl = pkg_slot[pkg]
            a, b = int(pkg_off[sl]), int(pkg_off[sl + 1])
            if b > a:
                agents_u.update(int(x) for x in ag[a:b])
                mh = int(pkg_minhop[sl])
                best = mh if best is None or mh < best else best
        va.extend(sorted(agents_u, key=lambda u: node_rank[u])); va_off.append(len(va))
        vmin.append(best if best is not None else 0)
    return dict(
        pkg_ids=pkg_ids, pkg_off=pkg_off, pkg_agents=ag.astype(np.int32), pkg_minhop=pkg_minhop,
        vuln_ids=vuln_ids, vuln_poff=np.asarray(vp_off, dtype=np.int64), vuln_pkgs=np.asarray(vp, dtype=np.int32),
        vuln_aoff=np.asarray(va_off, dtype=np.int64), vuln_agents=np.asarray(va, dtype=np.int32), vuln_minhop=np.asarray(vmin, dtype=np.int32),
    )

"""Lateral-movement paths over the legacy context graph, searched on the device.

Reference: ``/root/reference/src/agent_bom/context_graph.py`` — ``find_lateral_paths`` (:397-477) and
``_build_lateral_path`` (:480-593).  The order-sensitive queue search (100-path / 10 000-queue caps, per-path cycle
check, duplicate node sequences extended instead of recorded) runs in ``csrc/lateral.cuh``, one warp per source, so a
whole fleet is searched in one launch (the CLI and the REST route loop over every agent, ``cli/agents/__init__.py:1745``,
``api/routes/scan.py:701-705``).  Turning a found node sequence into a ``LateralPath`` — exposure lists, the 0..10
composite, the summary string — is label bookkeeping over at most 100 short paths per source and stays on the host.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Iterable

import numpy as np

from . import _lib
from .context_graph import kind_value

MAX_PATHS = 100                      # _MAX_PATHS (context_graph.py:393)
SEVERITY_RISK_SCORE = {"critical": 8.0, "high": 6.0, "medium": 4.0, "low": 2.0, "info": 0.5, "informational": 0.5, "none": 0.0, "unknown": 0.0}
_NODE_KIND_CODE = {"agent": 0, "server": 1, "credential": 2, "tool": 3, "vulnerability": 4, "iam_role": 5}
_GHOST = 255


@dataclass
class LateralPath:
    source: str
    target: str
    hops: list[str]
    edges: list[Any]
    composite_risk: float
    summary: str
    credential_exposure: list[str] = field(default_factory=list)
    tool_exposure: list[str] = field(default_factory=list)
    vuln_ids: list[str] = field(default_factory=list)


class LateralSearchOverflow(RuntimeError):
    """A source exhausted the search's safety valve; its result would be incomplete, so none is returned."""


class _Arrays:
    """``graph.adjacency`` as rows of (target, kind) and the per-node facts the target test reads."""

    def __init__(self, graph):
        nodes, adjacency = graph.nodes, graph.adjacency
        self.ids = list(nodes)
        index = {nid: i for i, nid in enumerate(self.ids)}
        rows = list(adjacency.items()) if hasattr(adjacency, "items") else []
        for nid, lst in rows:                                         # ids that only exist as adjacency keys / targets
            if nid not in index:
                index[nid] = len(self.ids)
                self.ids.append(nid)
            for e in lst:
                if e.target not in index:
                    index[e.target] = len(self.ids)
                    self.ids.append(e.target)
        n = len(self.ids)
        self.index = index
        self.kind_objs: list[Any] = []
        kind_code: dict[str, int] = {}
        counts = np.zeros(n + 1, dtype=np.int64)
        for nid, lst in rows:
            counts[index[nid] + 1] = len(lst)
        self.off = np.cumsum(counts)
        nbr = np.zeros(int(self.off[-1]), dtype=np.int32)
        ek = np.zeros(int(self.off[-1]), dtype=np.uint8)
        for nid, lst in rows:
            p = int(self.off[index[nid]])
            for e in lst:
                kv = kind_value(e.kind)
                code = kind_code.get(kv)
                if code is None:
                    code = kind_code[kv] = len(self.kind_objs)
                    self.kind_objs.append(e.kind)
                    if code > 15:
                        raise ValueError("more than 16 distinct edge kinds")
                nbr[p], ek[p] = index[e.target], code
                p += 1
        self.nbr, self.ek = nbr, ek
        self.keys: dict[Any, int] = {}
        self.nkind = np.full(n, _GHOST, dtype=np.uint8)
        self.nkey = np.full(n, -1, dtype=np.int32)
        for nid, node in nodes.items():
            i = index[nid]
            kv = kind_value(node.kind)
            self.nkind[i] = _NODE_KIND_CODE.get(kv, 6)
            if kv == "agent":
                self.nkey[i] = self.key_id(node.label)
            elif kv in ("credential", "tool"):
                owner = node.metadata.get("agent", "")
                if owner:
                    self.nkey[i] = self.key_id(owner)

    def key_id(self, value) -> int:
        return self.keys.setdefault(value, len(self.keys))


def _add(lst: list, item) -> bool:
    if item in lst:
        return False
    lst.append(item)
    return True


def build_lateral_path(graph, source_id: str, target_id: str, hops: list[str], edges: list) -> LateralPath:
    """Exposure lists, composite and summary of one found path (:480-593)."""
    nodes, adjacency = graph.nodes, graph.adjacency
    creds: list[str] = []
    tools: list[str] = []
    vulns: list[str] = []
    state = {"sev": 0.0, "exec": 0}

    def see_tool(node) -> None:
        if _add(tools, node.label) and "execute" in node.metadata.get("capabilities", []):
            state["exec"] += 1

    def see_vuln(node) -> None:
        if _add(vulns, node.label):
            state["sev"] = max(state["sev"], SEVERITY_RISK_SCORE.get(node.metadata.get("severity", ""), 0))

    def see_server_surface(server_id: str) -> None:
        """What a server exposes, provides and is vulnerable to, in adjacency order."""
        for e in adjacency.get(server_id, []):
            other = nodes.get(e.target)
            if not other:
                continue
            ek, ok = kind_value(e.kind), kind_value(other.kind)
            if ek == "exposes" and ok == "credential":
                _add(creds, other.label)
            elif ek == "provides" and ok == "tool":
                see_tool(other)
            elif ek == "vulnerable_to" and ok == "vulnerability":
                see_vuln(other)

    for nid in hops:
        node = nodes.get(nid)
        if not node:
            continue
        kv = kind_value(node.kind)
        if kv == "credential":
            _add(creds, node.label)
        elif kv == "tool":
            see_tool(node)
        elif kv == "vulnerability":
            see_vuln(node)
        elif kv == "server":
            see_server_surface(nid)
    # agent <-> agent sharing hops bypass the server node: fold in what the shared server / credential stands for (:538-572)
    for i, kind in enumerate(edges):
        if i >= len(hops) - 1:
            continue
        ek = kind_value(kind)
        if ek not in ("shares_server", "shares_credential"):
            continue
        link = next((e for e in adjacency.get(hops[i], []) if e.target == hops[i + 1] and kind_value(e.kind) == ek), None)
        if link is None:
            continue
        if ek == "shares_server":
            name = link.metadata.get("server", "")
            if name:
                for nid, node in nodes.items():
                    if kind_value(node.kind) == "server" and node.label == name:
                        see_server_surface(nid)
        else:
            name = link.metadata.get("credential", "")
            if name:
                _add(creds, name)
    composite = min(state["sev"] + len(creds) * 0.3 + state["exec"] * 0.2, 10.0)
    summary = " → ".join(nodes[n].label for n in hops if nodes.get(n))
    return LateralPath(source=source_id, target=target_id, hops=hops, edges=edges, composite_risk=round(composite, 1), summary=summary,
                       credential_exposure=creds, tool_exposure=tools, vuln_ids=vulns)



_ARRAYS_CACHE: dict[int, tuple[tuple, "_Arrays"]] = {}


def _arrays_for(graph) -> "_Arrays":
    """The array encoding of ``graph.adjacency`` is reused across calls on the same, unchanged graph (it was >99 % of a call:
    1.29 s of Python against 0.64 ms of device time for a 4 000-agent fleet); a change in node / edge / adjacency-row counts
    rebuilds it.  Entries die with their graph."""
    import weakref

    stamp = (len(graph.nodes), len(getattr(graph, "edges", ()) or ()), len(graph.adjacency), sum(map(len, graph.adjacency.values())) if len(graph.adjacency) < 4096 else -1)
    key = id(graph)
    hit = _ARRAYS_CACHE.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    arr = _Arrays(graph)
    _ARRAYS_CACHE[key] = (stamp, arr)
    try:
        weakref.finalize(graph, _ARRAYS_CACHE.pop, key, None)
    except TypeError:          # not weak-referenceable: do not keep it
        _ARRAYS_CACHE.pop(key, None)
    return arr


def search_many(graph, source_ids: Iterable[str], max_depth: int = 4, *, device: int = 0, max_pops: int = 0):
    """Raw device result per source: list of ``(hops, edge kinds)`` in discovery order."""
    lib = _lib.load()
    source_ids = list(source_ids)
    arr = _arrays_for(graph)
    src = np.full(len(source_ids), -1, dtype=np.int32)
    skey = np.zeros(len(source_ids), dtype=np.int32)
    for q, sid in enumerate(source_ids):
        node = graph.nodes.get(sid) if sid in graph.nodes else None
        if node is None:
            continue
        src[q] = arr.index[sid]
        skey[q] = arr.key_id(node.label if kind_value(node.kind) == "agent" else node.metadata.get("agent", ""))
    res = C.c_void_p()
    off64 = np.ascontiguousarray(arr.off, dtype=np.int64)
    _lib.check(lib.abb_lateral_paths_host(device, len(arr.ids), off64.ctypes.data, arr.nbr.ctypes.data, arr.ek.ctypes.data, arr.nkind.ctypes.data,
                                          arr.nkey.ctypes.data, len(source_ids), src.ctypes.data, skey.ctypes.data, int(max_depth), int(max_pops), C.byref(res)))
    try:
        nq = len(source_ids)
        W = int(lib.abb_lateral_result_width(res))
        R = W + 2
        off = np.ctypeslib.as_array(C.cast(lib.abb_lateral_result_off(res), C.POINTER(C.c_int64)), shape=(nq + 1,)).copy() if nq else np.zeros(1, np.int64)
        total = int(off[-1])
        rec = (np.ctypeslib.as_array(C.cast(lib.abb_lateral_result_records(res), C.POINTER(C.c_int32)), shape=(total * R,)).copy().reshape(total, R)
               if total else np.zeros((0, R), np.int32))
        flags = np.ctypeslib.as_array(C.cast(lib.abb_lateral_result_flags(res), C.POINTER(C.c_int32)), shape=(nq,)).copy() if nq else np.zeros(0, np.int32)
        ms = float(lib.abb_lateral_result_ms(res))
    finally:
        lib.abb_lateral_result_free(res)
    if flags.any():
        raise LateralSearchOverflow(f"search from {source_ids[int(np.flatnonzero(flags)[0])]!r} exceeded the safety valve")
    out = []
    for q in range(nq):
        found = []
        for row in rec[int(off[q]): int(off[q + 1])].tolist():
            length, kinds = row[0], row[1] & 0xFFFFFFFF
            hops = [arr.ids[i] for i in row[2: 2 + length]]
            edges = [arr.kind_objs[(kinds >> (4 * h)) & 15] for h in range(length - 1)]
            found.append((hops, edges))
        out.append(found)
    return out, ms


def find_lateral_paths_many(graph, source_ids: Iterable[str], max_depth: int = 4, *, device: int = 0) -> list[list[LateralPath]]:
    """``find_lateral_paths`` for many sources in one device launch; each list is sorted by composite risk, highest first."""
    source_ids = list(source_ids)
    raw, _ms = search_many(graph, source_ids, max_depth, device=device)
    out = []
    for sid, found in zip(source_ids, raw):
        paths = [build_lateral_path(graph, sid, hops[-1], hops, edges) for hops, edges in found]
        paths.sort(key=lambda p: p.composite_risk, reverse=True)
        out.append(paths[:MAX_PATHS])
    return out


def find_lateral_paths(graph, source_node_id: str, max_depth: int = 4, *, device: int = 0) -> list[LateralPath]:
    """Up to 100 lateral paths from one node, sorted by composite risk (drop-in for context_graph.py:397-477)."""
    if source_node_id not in graph.nodes:
        return []
    return find_lateral_paths_many(graph, [source_node_id], max_depth, device=device)[0]


__all__ = ["LateralPath", "LateralSearchOverflow", "build_lateral_path", "find_lateral_paths", "find_lateral_paths_many", "search_many"]

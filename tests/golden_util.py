"""Shared helpers for the parity tests: load golden fixtures, build oracle graphs, host-side path ranking."""

from __future__ import annotations

import gzip
import json
import sys
from functools import lru_cache
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import oracle as orc  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
ALL_FIXTURES = sorted(p.name[: -len(".json.gz")] for p in GOLDEN.glob("*.json.gz"))
SMALL_FIXTURES = [n for n in ALL_FIXTURES if n.startswith("kat_")]


@lru_cache(maxsize=None)
def load(name: str) -> dict:
    with gzip.open(GOLDEN / f"{name}.json.gz", "rb") as fh:
        return json.loads(fh.read())


def edge_arrays(doc: dict):
    e = np.asarray(doc["edges"], dtype=np.int64).reshape(-1, 4)
    return e[:, 0].astype(np.int32), e[:, 1].astype(np.int32), e[:, 2].astype(np.uint8), e[:, 3].astype(np.uint8)


def node_rank(ids) -> np.ndarray:
    """rank[i] = position of ids[i] in sorted(ids) — the order the reference's ``sorted()`` over id strings yields."""
    order = sorted(range(len(ids)), key=lambda i: ids[i])
    rank = np.empty(len(ids), dtype=np.int32)
    rank[np.asarray(order, dtype=np.int64)] = np.arange(len(ids), dtype=np.int32)
    return rank


@lru_cache(maxsize=None)
def oracle_graph(name: str):
    doc = load(name)
    src, dst, rel, flags = edge_arrays(doc)
    nt = np.asarray(doc["node_types"], dtype=np.uint8)
    return orc.build_csr(len(nt), src, dst, rel, flags, nt)


def seeded_graph(n_nodes: int, n_edges: int, seed: int, *, bidir_frac=0.15, nontrav_frac=0.1, ghost_frac=0.02, hub=True):
    """Random typed multigraph (unique (src,dst,rel) triples) used for differential CUDA-vs-oracle tests."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n_nodes, size=n_edges)
    dst = rng.integers(0, n_nodes, size=n_edges)
    if hub and n_nodes > 8:
        k = n_edges // 10
        src[:k] = rng.integers(0, 4, size=k)           # a few high-out-degree hubs
        dst[k: 2 * k] = rng.integers(4, 8, size=k)     # and high-in-degree hubs
    rel = rng.integers(0, 31, size=n_edges).astype(np.uint8)
    # over-represent the relationships the typed walks look for
    common = np.array([1, 2, 3, 4, 9, 13, 14, 7, 8, 26], dtype=np.uint8)
    pick = rng.random(n_edges) < 0.7
    rel[pick] = common[rng.integers(0, len(common), size=int(pick.sum()))]
    key = (src.astype(np.int64) * n_nodes + dst) * 32 + rel
    _, first = np.unique(key, return_index=True)
    first.sort()
    src, dst, rel = src[first], dst[first], rel[first]
    ne = len(src)
    flags = np.ones(ne, dtype=np.uint8)
    flags[rng.random(ne) < nontrav_frac] = 0
    flags[rng.random(ne) < bidir_frac] |= 2
    node_type = rng.integers(0, 24, size=n_nodes).astype(np.uint8)
    # make the typed pattern kinds frequent
    node_type[rng.random(n_nodes) < 0.6] = 0
    sel = rng.random(n_nodes)
    node_type[sel < 0.5] = np.array([0, 1, 2, 3, 8, 10, 9, 13, 17], dtype=np.uint8)[rng.integers(0, 9, size=int((sel < 0.5).sum()))]
    node_type[rng.random(n_nodes) < ghost_frac] = 255
    return src.astype(np.int32), dst.astype(np.int32), rel, flags, node_type


# ── host-side ranking of derived path rows (reference api/routes/graph.py:677-683, 762-786) ──

SEVERITY_RANK = {"critical": 5, "high": 4, "medium": 3, "low": 2, "info": 1, "informational": 1, "none": 0, "unknown": 0}


def node_risk_100(risk_score: float, severity: str) -> float:
    risk = float(risk_score or 0.0)
    if risk <= 10.0:
        risk *= 10.0
    if risk <= 0:
        risk = float(SEVERITY_RANK.get(str(severity or "").lower(), 0) * 20)
    return max(0.0, min(100.0, risk))


def rank_path_rows(doc: dict, rows, g) -> list[dict]:
    """Turn emission-order path rows into the reference's final sorted AttackPath dicts (index space)."""
    labels = doc["node_labels"]
    out = []
    for i in range(rows.hops.shape[0]):
        a, srv, vs, f = (int(x) for x in rows.hops[i])
        hops = [a, srv] + ([vs] if vs >= 0 else []) + [f]
        rels = [int(r) for r in rows.rels[i] if r >= 0]
        creds, tools = [], []
        for p in range(int(g.fwd_off[srv]), int(g.fwd_off[srv + 1])):
            m = int(g.fwd_meta[p])
            t = int(g.fwd_nbr[p])
            if m & 0x80 or g.node_type[t] == 255:
                continue
            if (m & 0x1F) == 4:
                creds.append(labels[t])
            elif (m & 0x1F) == 3:
                tools.append(labels[t])
        assert len(creds) == int(rows.ncred[i]) and len(tools) == int(rows.ntool[i])
        risk = node_risk_100(doc["node_risk"][f], doc["node_severity"][f])
        risk += min(10.0, len(creds) * 3.0)
        risk += min(10.0, len(tools) * 0.75)
        out.append({
            "hops": hops, "edges": rels, "risk": round(min(100.0, risk), 2), "creds": sorted(set(creds)), "tools": sorted(set(tools)),
            "vuln_ids": [labels[f] or doc["node_ids"][f]], "source": a, "target": f,
        })
    out.sort(key=lambda p: (p["risk"], len(p["hops"]), len(p["creds"]), len(p["tools"])), reverse=True)
    return out


def graph_from_fixture(doc: dict):
    """Our host container rebuilt from a golden fixture (string ids, labels, severities, risk scores)."""
    from agent_bom_b200.graph import EntityType, RelationshipType, UnifiedEdge, UnifiedGraph, UnifiedNode
    from agent_bom_b200.graph.schema import ENTITY_VALUES, RELATIONSHIP_VALUES

    g = UnifiedGraph(scan_id="golden", tenant_id="t")
    ids = doc["node_ids"]
    for i in range(doc["n_real"]):
        g.add_node(UnifiedNode(id=ids[i], entity_type=EntityType(ENTITY_VALUES[doc["node_types"][i]]), label=doc["node_labels"][i],
                               risk_score=doc["node_risk"][i], severity=doc["node_severity"][i]))
    for s, t, r, fl in doc["edges"]:
        g.add_edge(UnifiedEdge(source=ids[s], target=ids[t], relationship=RelationshipType(RELATIONSHIP_VALUES[r]) if r < 31 else "custom",
                               direction="bidirectional" if fl & 2 else "directed", traversable=bool(fl & 1)))
    return g

"""Effective-reach scoring: oracle vs the reference's answers (CPU), CUDA path vs both (GPU).

``tests/golden/context/effective_reach.json.gz`` holds ContextGraphs built by the unmodified reference (its own two test
fixtures — whose breakdowns equal the reference's checked-in snapshot file —, a seeded 40-agent fleet with shared
servers, hand-wired corner cases) and the reference's ``annotate_graph`` / ``compute`` output for each.
"""

from __future__ import annotations

import gzip
import json
from pathlib import Path

import numpy as np
import pytest

from agent_bom_b200.context_graph import ContextGraph, EdgeKind, GraphEdge, GraphNode, NodeKind
from agent_bom_b200.effective_reach import ReachScore, credential_tier, max_capability_weight

DOCS = json.loads(gzip.decompress((Path(__file__).parent / "golden" / "context" / "effective_reach.json.gz").read_bytes()))
IDS = [d["name"] for d in DOCS]


def rebuild(doc) -> ContextGraph:
    """The reference-built graph, re-made with this package's records; hand-wired adjacency entries are restored verbatim."""
    g = ContextGraph()
    for nid, kind, label, meta in doc["nodes"]:
        g.add_node(GraphNode(id=nid, kind=NodeKind(kind), label=label, metadata=dict(meta)))
    for s, t, k, meta in doc["edges"]:
        g.add_edge(GraphEdge(source=s, target=t, kind=EdgeKind(k), metadata=dict(meta)))
    for nid, lst in doc["adjacency"].items():
        have = [[e.source, e.target, e.kind.value] for e in g.adjacency.get(nid, [])]
        for s, t, k in lst[len(have):]:
            g.adjacency[nid].append(GraphEdge(source=s, target=t, kind=EdgeKind(k)))
    return g


@pytest.mark.parametrize("doc", DOCS, ids=IDS)
def test_container_mirrors_edges_like_the_reference(doc):
    g = rebuild(doc)
    got = {nid: [[e.source, e.target, e.kind.value] for e in lst] for nid, lst in g.adjacency.items() if lst}
    assert got == doc["adjacency"]
    assert len(g.edges) == len(doc["edges"])
    g.add_edge(GraphEdge(source=doc["edges"][0][0], target=doc["edges"][0][1], kind=EdgeKind(doc["edges"][0][2])))
    assert len(g.edges) == len(doc["edges"])                           # (source, target, kind) de-duplication


@pytest.mark.parametrize("doc", DOCS, ids=IDS)
def test_oracle_matches_the_reference(doc):
    from oracle import effective_reach_oracle as ero

    g = rebuild(doc)
    scores, edge_scores = ero.annotate(g)
    assert scores == doc["scores"] and list(scores) == list(doc["scores"])
    assert edge_scores == doc["edge_scores"]
    for nid, want in doc["extra"]:
        assert ero.score_node(g, g.nodes[nid]) == want


def test_formula_bands_and_tables():
    # tests/test_effective_reach.py:262-279 of the reference: band boundaries
    assert ReachScore(cvss=4.0, epss=0.05, is_kev=False, tool_capability=0.1, cred_visibility=0.1, agent_breadth=0).band == "green"
    assert ReachScore(cvss=8.0, epss=0.3, is_kev=False, tool_capability=0.4, cred_visibility=0.4, agent_breadth=1).band == "amber"
    assert ReachScore(cvss=9.0, epss=0.5, is_kev=True, tool_capability=0.5, cred_visibility=0.5, agent_breadth=2).band in ("red", "pulsing-red")
    top = ReachScore(cvss=9.8, epss=0.9, is_kev=True, tool_capability=1.0, cred_visibility=1.0, agent_breadth=5)
    assert top.band == "pulsing-red" and top.composite >= 90
    assert ReachScore(cvss=6.5, epss=0.02, is_kev=False, tool_capability=0.1, cred_visibility=0.0, agent_breadth=1).composite == 27.4     # reference snapshot
    assert ReachScore(cvss=50.0, epss=-3.0, is_kev=False, tool_capability=7.0, cred_visibility=-1.0, agent_breadth=99).composite == 80.0   # every input clamps
    from oracle import effective_reach_oracle as ero

    for name in ("AWS_ACCESS_KEY_ID", "github_token", " HOME ", "", None, "MY_SETTING", "SERVICE_PASSWORD", "DD_API_KEY", "path", "OAUTH_X", "apikey", "x"):
        assert credential_tier(name) == ero.tier(name)
    assert max_capability_weight(["read", "EXECUTE", "admin"]) == (1.0, "execute")
    assert max_capability_weight(["teleport"]) == (0.0, "") and max_capability_weight(None) == (0.0, "")
    assert max_capability_weight(["write", "write"]) == (0.65, "write")


# ── GPU ─────────────────────────────────────────────────────────────────────

@pytest.mark.gpu
@pytest.mark.parametrize("doc", DOCS, ids=IDS)
def test_device_scores_match_the_reference(doc):
    from agent_bom_b200.effective_reach import annotate_graph, compute, compute_many

    g = rebuild(doc)
    scores = annotate_graph(g)
    assert {nid: s.as_breakdown() for nid, s in scores.items()} == doc["scores"] and list(scores) == list(doc["scores"])
    assert [e.metadata.get("effective_reach_score") for e in g.edges] == doc["edge_scores"]
    for nid in scores:
        assert g.nodes[nid].metadata["effective_reach"] == doc["scores"][nid]
    for nid, want in doc["extra"]:
        assert compute(g.nodes[nid], g).as_breakdown() == want
    some = list(doc["scores"])[::3]
    assert {nid: s.as_breakdown() for nid, s in compute_many(g, some).items()} == {nid: doc["scores"][nid] for nid in some}
    assert compute_many(g, []) == {}


@pytest.mark.gpu
def test_device_scores_accept_the_oracle_on_a_larger_fleet():
    """Bigger than the goldens: 400 agents sharing 60 server names, 2 000 findings — CUDA path vs the per-finding oracle."""
    import random

    from agent_bom_b200.effective_reach import compute_many
    from oracle import effective_reach_oracle as ero

    rng = random.Random(7)
    g = ContextGraph()
    caps = [["read"], ["execute"], ["network", "write"], [], ["admin"], ["auth"], ["delete", "read"]]
    envs = ["AWS_KEY", "GITHUB_TOKEN", "HOME", "X_SECRET", "PLAIN", "DD_API_KEY", "REDIS_URL", "EDITOR"]
    names = [f"srv-{i:02d}" for i in range(60)]
    by_name: dict[str, list[str]] = {}
    servers = []
    for a in range(400):
        g.add_node(GraphNode(id=f"agent:{a}", kind=NodeKind.AGENT, label=f"agent-{a:03d}"))
        for name in rng.sample(names, rng.randint(1, 3)):
            sid = f"server:{a}:{name}"
            servers.append(sid)
            g.add_node(GraphNode(id=sid, kind=NodeKind.SERVER, label=name, metadata={"agent": f"agent-{a:03d}"}))
            g.add_edge(GraphEdge(source=f"agent:{a}", target=sid, kind=EdgeKind.USES))
            by_name.setdefault(name, []).append(f"agent:{a}")
            for t in range(rng.randint(0, 3)):
                tid = f"tool:{sid}:{t}"
                g.add_node(GraphNode(id=tid, kind=NodeKind.TOOL, label=f"tool-{rng.randint(0, 40)}", metadata={"capabilities": rng.choice(caps)}))
                g.add_edge(GraphEdge(source=sid, target=tid, kind=EdgeKind.PROVIDES))
            for key in rng.sample(envs, rng.randint(0, 3)):
                cid = f"cred:{key}"
                if cid not in g.nodes:
                    g.add_node(GraphNode(id=cid, kind=NodeKind.CREDENTIAL, label=key))
                g.add_edge(GraphEdge(source=sid, target=cid, kind=EdgeKind.EXPOSES))
    for name, agents in by_name.items():
        uniq = sorted(set(agents))
        for i, a1 in enumerate(uniq[:12]):
            for a2 in uniq[i + 1: 12]:
                g.add_edge(GraphEdge(source=a1, target=a2, kind=EdgeKind.SHARES_SERVER, metadata={"server": name}))
    for v in range(2000):
        vid = f"vuln:{v}"
        g.add_node(GraphNode(id=vid, kind=NodeKind.VULNERABILITY, label=f"CVE-{v}", metadata={"cvss_score": rng.choice([None, 4.4, 7.5, 9.8]),
                                                                                              "epss_score": rng.random(), "is_kev": rng.random() < 0.1}))
        for sid in rng.sample(servers, rng.randint(0, 5)):
            g.add_edge(GraphEdge(source=sid, target=vid, kind=EdgeKind.VULNERABLE_TO))
    got = compute_many(g)
    sample = list(got)[::25]
    for nid in sample:
        assert got[nid].as_breakdown() == ero.score_node(g, g.nodes[nid])
    assert len(got) == 2000


@pytest.mark.gpu
def test_group_union_primitive_against_numpy():
    from agent_bom_b200.engine import group_union

    rng = np.random.default_rng(11)
    n_members, n_groups = 5000, 3000
    item_counts = rng.integers(0, 40, n_members)
    item_counts[rng.integers(0, n_members, 20)] = 3000                      # a few very long lists
    ioff = np.concatenate([[0], np.cumsum(item_counts)]).astype(np.int64)
    items = rng.integers(0, 1 << 30, int(ioff[-1]), dtype=np.int64).astype(np.int32)
    items[rng.integers(0, items.size, items.size // 3)] = 7                   # plenty of duplicates
    mcounts = rng.integers(0, 9, n_groups)
    mcounts[5] = 400
    moff = np.concatenate([[0], np.cumsum(mcounts)]).astype(np.int64)
    members = rng.integers(0, n_members, int(moff[-1])).astype(np.int32)
    w0 = rng.integers(0, 8, n_members).astype(np.uint8)
    w1 = rng.integers(0, 4, n_members).astype(np.uint8)
    off, out, g0, g1, ms = group_union(moff, members, ioff, items, w0, w1)
    assert ms >= 0.0 and off[0] == 0 and off.shape == (n_groups + 1,)
    for gi in range(n_groups):
        ms_ = members[moff[gi]: moff[gi + 1]]
        want = np.unique(np.concatenate([items[ioff[m]: ioff[m + 1]] for m in ms_])) if len(ms_) else np.zeros(0, np.int32)
        assert np.array_equal(out[off[gi]: off[gi + 1]], want), gi
        assert g0[gi] == (w0[ms_].max() if len(ms_) else 0) and g1[gi] == (w1[ms_].max() if len(ms_) else 0)
    # degenerate shapes
    off, out, g0, g1, _ = group_union([0], [], [0], [])
    assert off.tolist() == [0] and out.size == 0 and g0.size == 0
    off, out, g0, g1, _ = group_union([0, 0, 2], [0, 0], [0, 3], [5, 5, 1])
    assert off.tolist() == [0, 0, 2] and out.tolist() == [1, 5] and g0.tolist() == [0, 0]
    with pytest.raises(Exception):
        group_union([0, 1], [3], [0, 0], [])                                   # member index out of range

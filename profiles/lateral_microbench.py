"""Device timing of the lateral-path search on a synthetic fleet, with the CPU oracle timed on a sample.

    python profiles/lateral_microbench.py [agents] > profiles/r01_lateral_microbench.json

Fleet: every agent uses 2 private servers and 1–2 of `agents/25` shared server names (SHARES_SERVER cliques of ≈ 40–50 agents),
each server exposes credentials and provides tools.  All agents are searched in one launch (max_depth 4), as the reference's CLI /
REST callers do one agent at a time.  `device_ms` is the kernel's CUDA-event time; `oracle_ms_per_source` is the pure-Python
restatement of the reference's queue loop (oracle/lateral_oracle.py) on 24 evenly spaced sources — its results are compared too."""

from __future__ import annotations

import json
import random
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from agent_bom_b200.context_graph import ContextGraph, EdgeKind, GraphEdge, GraphNode, NodeKind  # noqa: E402
from agent_bom_b200.lateral import search_many  # noqa: E402
from oracle import lateral_oracle as lo  # noqa: E402


def fleet(n_agents: int) -> ContextGraph:
    rng = random.Random(11)
    g = ContextGraph()
    shared = [f"shared-{i:03d}" for i in range(max(2, n_agents // 25))]
    users: dict[str, list[str]] = {}
    for a in range(n_agents):
        name = f"agent-{a:05d}"
        g.add_node(GraphNode(id=f"agent:{name}", kind=NodeKind.AGENT, label=name))
        for sname in [f"own-{a}-0", f"own-{a}-1"] + rng.sample(shared, rng.randint(1, 2)):
            sid = f"server:{name}:{sname}"
            g.add_node(GraphNode(id=sid, kind=NodeKind.SERVER, label=sname, metadata={"agent": name}))
            g.add_edge(GraphEdge(source=f"agent:{name}", target=sid, kind=EdgeKind.USES))
            if sname in shared:
                users.setdefault(sname, []).append(name)
            for t in range(2):
                tid = f"tool:{sid}:{t}"
                g.add_node(GraphNode(id=tid, kind=NodeKind.TOOL, label=f"tool-{t}", metadata={"agent": name, "capabilities": ["execute"] if t == 0 else ["read"]}))
                g.add_edge(GraphEdge(source=sid, target=tid, kind=EdgeKind.PROVIDES))
            cid = f"cred:K{rng.randrange(n_agents // 10 + 1)}"
            if cid not in g.nodes:
                g.add_node(GraphNode(id=cid, kind=NodeKind.CREDENTIAL, label=cid[5:], metadata={"servers": []}))
            g.add_edge(GraphEdge(source=sid, target=cid, kind=EdgeKind.EXPOSES))
    for sname, names in users.items():
        uniq = sorted(set(names))
        for i, a1 in enumerate(uniq):
            for a2 in uniq[i + 1:]:
                g.add_edge(GraphEdge(source=f"agent:{a1}", target=f"agent:{a2}", kind=EdgeKind.SHARES_SERVER, metadata={"server": sname}))
    return g


def main() -> None:
    n_agents = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    g = fleet(n_agents)
    sources = [nid for nid, n in g.nodes.items() if n.kind == NodeKind.AGENT]
    search_many(g, sources[:8], 4)                                   # warm-up (context, module load)
    t = time.perf_counter()
    found, ms = search_many(g, sources, 4)
    wall = (time.perf_counter() - t) * 1e3
    sample = sources[:: max(1, len(sources) // 24)]
    t = time.perf_counter()
    want = [lo.search(g, s, 4) for s in sample]
    oracle_ms = (time.perf_counter() - t) * 1e3 / len(sample)
    same = all([(h, [k.value for k in ks]) for h, ks in found[sources.index(s)]] == [(h, [k.value for k in ks]) for h, ks in w] for s, w in zip(sample, want))
    print(json.dumps({
        "kernel": "lateral_search_kernel (one warp per source)", "agents": n_agents, "nodes": len(g.nodes), "edges": len(g.edges),
        "adjacency_entries": sum(len(v) for v in g.adjacency.values()), "sources": len(sources), "max_depth": 4,
        "paths_found": sum(len(f) for f in found), "device_ms": ms, "sources_per_s_device": len(sources) / (ms * 1e-3),
        "call_ms_with_encoding_and_copies": wall, "oracle_ms_per_source": oracle_ms, "oracle_sources": len(sample),
        "device_vs_oracle_per_source": oracle_ms / (ms / len(sources)), "sample_identical_to_oracle": same}))


if __name__ == "__main__":
    main()

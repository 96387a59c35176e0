"""CPU oracle for lateral-path search — plain Python, small cases only.  TEST INFRASTRUCTURE ONLY.

Restates ``/root/reference/src/agent_bom/context_graph.py``: the queue search of ``find_lateral_paths`` (:397-477)
and the scoring of ``_build_lateral_path`` (:480-593).  Pinned against the unmodified reference's output in
``tests/golden/context/lateral.json.gz`` (``oracle/make_golden.py --lateral-only``); the tests then hold the CUDA
search (``agent_bom_b200.lateral``) against it on graphs larger than the goldens.
"""

from __future__ import annotations

from collections import deque

SEV = {"critical": 8.0, "high": 6.0, "medium": 4.0, "low": 2.0, "info": 0.5, "informational": 0.5, "none": 0.0, "unknown": 0.0}


def _k(kind) -> str:
    return getattr(kind, "value", kind)


def search(graph, source_id, max_depth=4, max_paths=100, max_queue=10_000):
    """[(hops, kinds)] in discovery order."""
    if source_id not in graph.nodes:
        return []
    start = graph.nodes[source_id]
    me = start.label if _k(start.kind) == "agent" else start.metadata.get("agent", "")
    todo = deque([((source_id,), ())])
    recorded, seen = [], set()
    while todo and len(recorded) < max_paths:
        hops, kinds = todo.popleft()
        if len(hops) > max_depth + 1:
            continue
        tip = hops[-1]
        if len(hops) > 1:
            node = graph.nodes.get(tip)
            hit = False
            if node:
                nk = _k(node.kind)
                if nk == "agent":
                    hit = node.label != me
                elif nk in ("credential", "tool"):
                    owner = node.metadata.get("agent", "")
                    hit = bool(owner) and owner != me
            if hit and hops not in seen:
                seen.add(hops)
                recorded.append((list(hops), list(kinds)))
                continue
        if len(todo) >= max_queue:
            continue
        for e in graph.adjacency.get(tip, []):
            if e.target not in hops:
                todo.append((hops + (e.target,), kinds + (e.kind,)))
    return recorded


def describe(graph, hops, kinds):
    """(composite, summary, credentials, tools, vulnerabilities) of one path."""
    creds, tools, vulns = [], [], []
    box = [0.0, 0]

    def surface(sid):
        for e in graph.adjacency.get(sid, []):
            o = graph.nodes.get(e.target)
            if not o:
                continue
            pair = (_k(e.kind), _k(o.kind))
            if pair == ("exposes", "credential"):
                note(o)
            elif pair in (("provides", "tool"), ("vulnerable_to", "vulnerability")):
                note(o)

    def note(o):
        k = _k(o.kind)
        if k == "credential":
            if o.label not in creds:
                creds.append(o.label)
        elif k == "tool":
            if o.label not in tools:
                tools.append(o.label)
                if "execute" in o.metadata.get("capabilities", []):
                    box[1] += 1
        elif k == "vulnerability":
            if o.label not in vulns:
                vulns.append(o.label)
                box[0] = max(box[0], SEV.get(o.metadata.get("severity", ""), 0))

    for h in hops:
        n = graph.nodes.get(h)
        if not n:
            continue
        if _k(n.kind) == "server":
            surface(h)
        else:
            note(n)
    for i, kind in enumerate(kinds):
        if i >= len(hops) - 1 or _k(kind) not in ("shares_server", "shares_credential"):
            continue
        for e in graph.adjacency.get(hops[i], []):
            if e.target == hops[i + 1] and _k(e.kind) == _k(kind):
                if _k(kind) == "shares_server":
                    name = e.metadata.get("server", "")
                    if name:
                        for sid, n in graph.nodes.items():
                            if _k(n.kind) == "server" and n.label == name:
                                surface(sid)
                else:
                    name = e.metadata.get("credential", "")
                    if name and name not in creds:
                        creds.append(name)
                break
    total = min(box[0] + len(creds) * 0.3 + box[1] * 0.2, 10.0)
    return round(total, 1), " → ".join(graph.nodes[h].label for h in hops if graph.nodes.get(h)), creds, tools, vulns


def find(graph, source_id, max_depth=4):
    """The reference's final answer as plain dicts, sorted by composite risk (stable)."""
    out = []
    for hops, kinds in search(graph, source_id, max_depth):
        risk, summary, creds, tools, vulns = describe(graph, hops, kinds)
        out.append({"source": source_id, "target": hops[-1], "hops": hops, "edges": [_k(k) for k in kinds], "composite_risk": risk, "summary": summary,
                    "credential_exposure": creds, "tool_exposure": tools, "vuln_ids": vulns})
    out.sort(key=lambda p: p["composite_risk"], reverse=True)
    return out[:100]

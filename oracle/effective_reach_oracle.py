"""CPU oracle for effective-reach scoring — a plain-Python restatement, sets and loops, small cases only.

TEST INFRASTRUCTURE ONLY.  Nothing in the product imports this; tests check the CUDA path
(``agent_bom_b200.effective_reach``) against it and it is itself pinned against the unmodified reference's answers in
``tests/golden/context/effective_reach.json.gz`` (``oracle/make_golden.py --effective-reach-only``), which include the
reference's own snapshot fixture ``tests/fixtures/effective_reach_snapshots.json``.

Follows ``/root/reference/src/agent_bom/effective_reach.py``: weights :61-69, ``_credential_tier`` :144-169,
``ReachScore.composite`` / ``band`` / ``as_breakdown`` :192-262, ``_affected_servers`` :265-279, per-server lookups
:281-353, ``compute`` :372-426, ``annotate_graph`` :429-465.  Written per finding, exactly as expensive as the
reference (every finding rescans the edge list) — that cost is the point of the device version.
"""

from __future__ import annotations

WEIGHTS = {"read": 0.10, "network": 0.40, "auth": 0.55, "write": 0.65, "delete": 0.75, "admin": 0.85, "execute": 1.00}
CLOUD = ("AWS_", "AMAZON_", "GCP_", "GOOGLE_", "AZURE_", "MS_", "OPENAI_", "ANTHROPIC_", "CLAUDE_", "GEMINI_", "BEDROCK_", "VERTEX_", "DATABRICKS_",
         "SNOWFLAKE_", "STRIPE_", "TWILIO_", "DD_", "PAGERDUTY_")
PROJECT = ("GITHUB_", "GITLAB_", "BITBUCKET_", "NPM_", "PYPI_", "DOCKER_", "GHCR_", "DATABASE_", "DB_", "POSTGRES_", "MYSQL_", "REDIS_", "MONGODB_",
           "JIRA_", "SLACK_", "NOTION_", "LINEAR_", "OAUTH_")
SHELL = {"HOME", "USER", "USERNAME", "LOGNAME", "PWD", "OLDPWD", "SHELL", "TERM", "LANG", "LC_ALL", "LC_CTYPE", "PATH", "TMPDIR", "DISPLAY", "EDITOR"}


def _k(kind) -> str:
    return getattr(kind, "value", kind)


def tier(name) -> float:
    key = (name or "").strip().upper()
    if not key:
        return 0.0
    for p in CLOUD:
        if key.startswith(p):
            return 1.0
    for p in PROJECT:
        if key.startswith(p):
            return 0.55
    if key in SHELL:
        return 0.10
    for word in ("TOKEN", "SECRET", "KEY", "PASSWORD", "API"):
        if word in key:
            return 0.55
    return 0.10


def composite(cvss, epss, kev, tool, cred, breadth) -> float:
    c = max(0.0, min(cvss, 10.0))
    e = max(0.0, min(epss, 1.0))
    t = max(0.0, min(tool, 1.0))
    r = max(0.0, min(cred, 1.0))
    b = max(0, min(breadth, 5))
    total = (c / 10.0) * 30.0 + e * 20.0 + (40.0 if kev else 0.0) + t * 25.0 + r * 20.0 + b * 5.0
    return round(max(0.0, min(total, 100.0)), 2)


def band(score: float) -> str:
    if score >= 90.0:
        return "pulsing-red"
    if score > 70.0:
        return "red"
    if score > 30.0:
        return "amber"
    return "green"


def score_node(graph, node) -> dict:
    """Breakdown dict of one node."""
    nodes, edges, adjacency = graph.nodes, graph.edges, graph.adjacency
    best_tool = best_cred = 0.0
    tool_names, cred_names, agent_names = set(), set(), set()
    for sid in sorted(e.source for e in edges if _k(e.kind) == "vulnerable_to" and e.target == node.id):
        for e in adjacency.get(sid, []):
            other = nodes.get(e.target)
            if other is None:
                continue
            if _k(e.kind) == "provides" and _k(other.kind) == "tool":
                w, named = 0.0, False
                for cap in other.metadata.get("capabilities") or []:
                    cw = WEIGHTS.get(str(cap).lower(), 0.0)
                    if cw > w:
                        w, named = cw, True
                best_tool = max(best_tool, w)
                if named:
                    tool_names.add(other.label)
            elif _k(e.kind) == "exposes" and _k(other.kind) == "credential":
                best_cred = max(best_cred, tier(other.label))
                cred_names.add(other.label)
        server = nodes.get(sid)
        if not server:
            continue
        if server.metadata.get("agent"):
            agent_names.add(str(server.metadata["agent"]))
        for e in edges:
            if _k(e.kind) == "uses" and e.target == sid:
                a = nodes.get(e.source)
                if a and _k(a.kind) == "agent":
                    agent_names.add(a.label)
            if _k(e.kind) == "shares_server" and e.metadata.get("server") == server.label:
                for end in (e.source, e.target):
                    a = nodes.get(end)
                    if a and _k(a.kind) == "agent":
                        agent_names.add(a.label)
        for e in adjacency.get(sid, []):
            if _k(e.kind) == "uses":
                a = nodes.get(e.target)
                if a and _k(a.kind) == "agent":
                    agent_names.add(a.label)
    cvss = float(node.metadata.get("cvss_score") or 0.0)
    epss = float(node.metadata.get("epss_score") or 0.0)
    kev = bool(node.metadata.get("is_kev"))
    total = composite(cvss, epss, kev, best_tool, best_cred, len(agent_names))
    return {"cvss": round(cvss, 2), "epss": round(epss, 4), "is_kev": kev, "tool_capability": round(best_tool, 3), "cred_visibility": round(best_cred, 3),
            "agent_breadth": len(agent_names), "reachable_tools": sorted(tool_names), "reachable_creds": sorted(cred_names),
            "reachable_agents": sorted(agent_names), "composite": total, "band": band(total)}


def annotate(graph):
    """(scores by vulnerability id in node order, per-edge inherited composite or None)."""
    scores = {nid: score_node(graph, n) for nid, n in graph.nodes.items() if _k(n.kind) == "vulnerability"}
    edge_scores = []
    for e in graph.edges:
        cands = [scores[x]["composite"] for x in (e.source, e.target) if x in scores]
        edge_scores.append(max(cands) if cands else None)
    return scores, edge_scores

"""Effective-reach scoring with the per-vulnerability reduction on the device.

Reference: ``/root/reference/src/agent_bom/effective_reach.py`` — ``ReachScore`` (:173-262), ``compute`` (:372-426),
``annotate_graph`` (:429-465), the capability weights (:61-69) and the credential-tier heuristic (:144-169).

Split of work
-------------
* Host, once per graph, O(|edges| + |adjacency|): label tables (sorted, so integer order == string order), one byte
  weight per tool / credential label, and three per-server item lists (tool labels that carry a capability, credential
  labels, agent labels that can pivot through the server — ``_tools_for_server`` / ``_creds_for_server`` /
  ``_agents_for_server``, :281-353).  The reference rescans ``graph.edges`` for every vulnerability and every server
  (O(V·E)); here one pass builds the indexes.
* Device, for all vulnerabilities at once (``abb_group_union_host``): per vulnerability the de-duplicated sorted union
  of its servers' item lists and the maxima of the two weights.
* Host: the float formula on the reduced inputs, with the reference's expression order and Python ``round``
  (bit-identical composites; numpy float64 evaluates the same IEEE operations).

There is no CPU fallback: without the CUDA library the scorer raises ``EngineUnavailable``.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Iterable, Literal

import numpy as np

from .context_graph import kind_value

# 0..1 weight of what a reachable tool can do; several tools combine by max (effective_reach.py:61-69)
CAPABILITY_WEIGHT: dict[str, float] = {"read": 0.10, "network": 0.40, "auth": 0.55, "write": 0.65, "delete": 0.75, "admin": 0.85, "execute": 1.00}
_TOOL_LEVELS = (0.0,) + tuple(sorted(set(CAPABILITY_WEIGHT.values())))          # byte code -> weight, ascending so max(code) == max(weight)
_CRED_LEVELS = (0.0, 0.10, 0.55, 1.0)

# prefixes of cloud / SaaS API credentials (tier 3), project-scoped credentials (tier 2) and shell defaults (tier 1) — :82-141
_CLOUD_PREFIXES = ("AWS_", "AMAZON_", "GCP_", "GOOGLE_", "AZURE_", "MS_", "OPENAI_", "ANTHROPIC_", "CLAUDE_", "GEMINI_", "BEDROCK_", "VERTEX_",
                   "DATABRICKS_", "SNOWFLAKE_", "STRIPE_", "TWILIO_", "DD_", "PAGERDUTY_")
_PROJECT_PREFIXES = ("GITHUB_", "GITLAB_", "BITBUCKET_", "NPM_", "PYPI_", "DOCKER_", "GHCR_", "DATABASE_", "DB_", "POSTGRES_", "MYSQL_", "REDIS_",
                     "MONGODB_", "JIRA_", "SLACK_", "NOTION_", "LINEAR_", "OAUTH_")
_HOME_KEYS = frozenset({"HOME", "USER", "USERNAME", "LOGNAME", "PWD", "OLDPWD", "SHELL", "TERM", "LANG", "LC_ALL", "LC_CTYPE", "PATH", "TMPDIR",
                        "DISPLAY", "EDITOR"})
_SECRET_WORDS = ("TOKEN", "SECRET", "KEY", "PASSWORD", "API")

_CAT_SHIFT = 29            # item = category << 29 | label rank; categories sort as tools < credentials < agents
_CAT_TOOL, _CAT_CRED, _CAT_AGENT = 0, 1, 2


def credential_tier(env_key: str) -> float:
    """Visibility weight of an env-var name (:144-169): cloud prefix 1.0, project prefix 0.55, shell default 0.10,
    otherwise 0.55 when the name looks like a secret and 0.10 when it does not; blank names weigh nothing."""
    key = (env_key or "").strip().upper()
    if not key:
        return 0.0
    if key.startswith(_CLOUD_PREFIXES):
        return 1.0
    if key.startswith(_PROJECT_PREFIXES):
        return 0.55
    if key in _HOME_KEYS:
        return 0.10
    return 0.55 if any(word in key for word in _SECRET_WORDS) else 0.10


def max_capability_weight(capabilities) -> tuple[float, str]:
    """Strongest capability of a tool and its label; the first one wins ties (:356-366)."""
    best, label = 0.0, ""
    for cap in capabilities or []:
        name = str(cap).lower()
        w = CAPABILITY_WEIGHT.get(name, 0.0)
        if w > best:
            best, label = w, name
    return best, label


@dataclass(frozen=True)
class ReachScore:
    """Inputs and result of the composite for one finding (:173-262)."""

    cvss: float
    epss: float
    is_kev: bool
    tool_capability: float
    cred_visibility: float
    agent_breadth: int
    reachable_tools: tuple[str, ...] = field(default=())
    reachable_creds: tuple[str, ...] = field(default=())
    reachable_agents: tuple[str, ...] = field(default=())

    @property
    def composite(self) -> float:
        """0..100: 30 for CVSS, 20 for EPSS, +40 KEV, 25 tool capability, 20 credential visibility, 5 per agent up to 5 (:192-232)."""
        cvss = max(0.0, min(self.cvss, 10.0))
        epss = max(0.0, min(self.epss, 1.0))
        tool = max(0.0, min(self.tool_capability, 1.0))
        cred = max(0.0, min(self.cred_visibility, 1.0))
        breadth = max(0, min(self.agent_breadth, 5))
        score = (cvss / 10.0) * 30.0 + epss * 20.0 + (40.0 if self.is_kev else 0.0) + tool * 25.0 + cred * 20.0 + breadth * 5.0
        return round(max(0.0, min(score, 100.0)), 2)

    @property
    def band(self) -> Literal["green", "amber", "red", "pulsing-red"]:
        c = self.composite
        return "pulsing-red" if c >= 90.0 else "red" if c > 70.0 else "amber" if c > 30.0 else "green"

    def as_breakdown(self) -> dict[str, object]:
        return {"cvss": round(self.cvss, 2), "epss": round(self.epss, 4), "is_kev": self.is_kev, "tool_capability": round(self.tool_capability, 3),
                "cred_visibility": round(self.cred_visibility, 3), "agent_breadth": self.agent_breadth, "reachable_tools": list(self.reachable_tools),
                "reachable_creds": list(self.reachable_creds), "reachable_agents": list(self.reachable_agents), "composite": self.composite,
                "band": self.band}


class _Encoded:
    """Array form of a context graph for the device reduction."""

    def __init__(self, graph, finding_ids: list[str]):
        nodes = graph.nodes
        edges = graph.edges
        adjacency = graph.adjacency
        want = {fid: i for i, fid in enumerate(finding_ids)}
        # one pass over graph.edges: servers per finding (:265-279), USES sources per server and SHARES_SERVER endpoints per server name (:318-334)
        servers_of: list[list[str]] = [[] for _ in finding_ids]
        uses_agents: dict[str, list[str]] = {}
        shared_by_name: dict[Any, list[str]] = {}
        for e in edges:
            k = kind_value(e.kind)
            if k == "vulnerable_to":
                gi = want.get(e.target)
                if gi is not None:
                    servers_of[gi].append(e.source)
            elif k == "uses":
                agent = nodes.get(e.source)
                if agent is not None and kind_value(agent.kind) == "agent":
                    uses_agents.setdefault(e.target, []).append(agent.label)
            elif k == "shares_server":
                try:
                    bucket = shared_by_name.setdefault(e.metadata.get("server"), [])
                except TypeError:          # an unhashable value can never equal a server label
                    continue
                for end in (e.source, e.target):
                    agent = nodes.get(end)
                    if agent is not None and kind_value(agent.kind) == "agent":
                        bucket.append(agent.label)
        member_ids: list[str] = []
        member_index: dict[str, int] = {}
        moff = [0]
        members: list[int] = []
        for lst in servers_of:
            for sid in sorted(lst):
                mi = member_index.get(sid)
                if mi is None:
                    mi = member_index[sid] = len(member_ids)
                    member_ids.append(sid)
                members.append(mi)
            moff.append(len(members))
        # per member: labels per category and the two weights
        per_member: list[tuple[list[str], list[str], list[str]]] = []
        w_tool = np.zeros(len(member_ids), dtype=np.uint8)
        w_cred = np.zeros(len(member_ids), dtype=np.uint8)
        tool_code = {w: i for i, w in enumerate(_TOOL_LEVELS)}
        cred_code = {w: i for i, w in enumerate(_CRED_LEVELS)}
        for mi, sid in enumerate(member_ids):
            tools: list[str] = []
            creds: list[str] = []
            agents: list[str] = []
            adj = adjacency.get(sid, []) if hasattr(adjacency, "get") else []
            for e in adj:
                k = kind_value(e.kind)
                other = nodes.get(e.target)
                if other is None:
                    continue
                ok = kind_value(other.kind)
                if k == "provides" and ok == "tool":
                    w, label = max_capability_weight(other.metadata.get("capabilities"))
                    w_tool[mi] = max(int(w_tool[mi]), tool_code[w])
                    if label:
                        tools.append(other.label)
                elif k == "exposes" and ok == "credential":
                    w_cred[mi] = max(int(w_cred[mi]), cred_code[credential_tier(other.label)])
                    creds.append(other.label)
                elif k == "uses" and ok == "agent":
                    agents.append(other.label)                      # callers that wired server -> agent by hand (:347-352)
            server = nodes.get(sid)
            if server is not None:
                owner = server.metadata.get("agent")
                if owner:
                    agents.append(str(owner))
                agents.extend(uses_agents.get(sid, ()))
                try:
                    agents.extend(shared_by_name.get(server.label, ()))
                except TypeError:
                    pass
            else:
                agents = []                                          # an unknown server has no pivoting agents (:311-313)
            per_member.append((tools, creds, agents))
        self.tables = []
        for cat in range(3):
            self.tables.append(sorted({lab for row in per_member for lab in row[cat]}, key=_label_key))
        rank = [{lab: i for i, lab in enumerate(t)} for t in self.tables]
        if any(len(t) >= (1 << _CAT_SHIFT) for t in self.tables):
            raise ValueError("more than 2^29 distinct labels in one category")
        ioff = [0]
        items: list[int] = []
        for row in per_member:
            for cat in range(3):
                r = rank[cat]
                base = cat << _CAT_SHIFT
                items.extend(base | r[lab] for lab in row[cat])
            ioff.append(len(items))
        self.member_off = np.asarray(moff, dtype=np.int64)
        self.members = np.asarray(members, dtype=np.int32)
        self.item_off = np.asarray(ioff, dtype=np.int64)
        self.items = np.asarray(items, dtype=np.int32)
        self.w_tool, self.w_cred = w_tool, w_cred


def _label_key(label):
    """Labels sort as the reference sorts them (plain ``sorted`` of strings); non-strings are ordered by their text."""
    return label if isinstance(label, str) else str(label)


def compute_many(graph, finding_ids: Iterable[str] | None = None, *, device: int = 0) -> dict[str, ReachScore]:
    """Scores for the given finding nodes (default: every vulnerability node, in ``graph.nodes`` order, :447-453)."""
    from .engine import group_union

    if finding_ids is None:
        finding_ids = [nid for nid, n in graph.nodes.items() if kind_value(n.kind) == "vulnerability"]
    finding_ids = list(finding_ids)
    enc = _Encoded(graph, finding_ids)
    off, items, g_tool, g_cred, _ms = group_union(enc.member_off, enc.members, enc.item_off, enc.items, enc.w_tool, enc.w_cred, device=device)
    cats = items >> _CAT_SHIFT
    ranks = (items & ((1 << _CAT_SHIFT) - 1)).tolist()
    cat_list = cats.tolist()
    out: dict[str, ReachScore] = {}
    tools_t, creds_t, agents_t = enc.tables
    for gi, fid in enumerate(finding_ids):
        node = graph.nodes[fid]
        a, b = int(off[gi]), int(off[gi + 1])
        tools: list[str] = []
        creds: list[str] = []
        agents: list[str] = []
        for c, r in zip(cat_list[a:b], ranks[a:b]):
            (tools if c == _CAT_TOOL else creds if c == _CAT_CRED else agents).append((tools_t, creds_t, agents_t)[c][r])
        meta = node.metadata
        out[fid] = ReachScore(
            cvss=float(meta.get("cvss_score") or 0.0), epss=float(meta.get("epss_score") or 0.0), is_kev=bool(meta.get("is_kev")),
            tool_capability=_TOOL_LEVELS[int(g_tool[gi])], cred_visibility=_CRED_LEVELS[int(g_cred[gi])], agent_breadth=len(agents),
            reachable_tools=tuple(tools), reachable_creds=tuple(creds), reachable_agents=tuple(agents))
    return out


def compute(node, graph, *, device: int = 0) -> ReachScore:
    """Score of one node (any kind — a node nothing is VULNERABLE_TO gets the degenerate score, :375-379)."""
    if graph.nodes.get(node.id) is not node:
        shadow = _Shadow(graph, node)
        return compute_many(shadow, [node.id], device=device)[node.id]
    return compute_many(graph, [node.id], device=device)[node.id]


class _Shadow:
    """A view of ``graph`` in which ``node`` answers for its id (``compute`` may be handed a detached node)."""

    def __init__(self, graph, node):
        self.nodes = dict(graph.nodes)
        self.nodes[node.id] = node
        self.edges, self.adjacency = graph.edges, graph.adjacency


def annotate_graph(graph, *, device: int = 0) -> dict[str, ReachScore]:
    """Score every vulnerability node, store the breakdown on it, and let each edge carry the higher composite of its
    scored endpoints (:429-465)."""
    scores = compute_many(graph, device=device)
    for nid, score in scores.items():
        graph.nodes[nid].metadata["effective_reach"] = score.as_breakdown()
    if not scores:
        return scores
    composite = {nid: s.composite for nid, s in scores.items()}
    for edge in graph.edges:
        a, b = composite.get(edge.source), composite.get(edge.target)
        if a is None and b is None:
            continue
        edge.metadata["effective_reach_score"] = a if b is None else b if a is None else max(a, b)
    return scores


__all__ = ["CAPABILITY_WEIGHT", "ReachScore", "annotate_graph", "compute", "compute_many", "credential_tier", "max_capability_weight"]

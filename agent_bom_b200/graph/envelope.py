"""ExposurePath wire envelopes for attack paths — the step right after the path.

Produces, for ``AttackPath`` records made on the device (``graph/exposure.py``) or loaded from a snapshot, the two JSON
objects the reference emits for them:

* REST — ``exposure_path`` inside each element of ``attack_paths``
  (behaviour of ``/root/reference/src/agent_bom/api/routes/graph.py:506-670`` with its helpers ``:275-291``, ``:484-503``);
* MCP — the ``exposure_paths`` tool payload (behaviour of ``/root/reference/src/agent_bom/mcp_tools/graph.py:16-100``).

The wire contract (key names, key order, fallbacks) is fixed by the reference; the construction here is table-driven and
indexed: the reference rebuilds an O(|E|) ``(source, target) → edge`` map for EVERY path, here one ``EdgeIndex`` per graph answers
"first edge of a hop pair" (REST; a bidirectional edge also answers for the reversed pair), "all edges of a hop pair" and
"edges with this id" (MCP).  Pinned object-for-object by ``tests/test_envelope.py`` on envelopes the unmodified reference
produced (``tests/golden/envelope/``, ``oracle/make_golden.py``).
"""

from __future__ import annotations

from typing import Any, Iterable

from .schema import SEVERITY_RANK, enum_value

# entity type -> role of a hop in the cockpit
_ROLE_GROUPS = (
    ("finding", ("vulnerability", "misconfiguration")),
    ("package", ("package",)),
    ("server", ("server", "container", "cloud_resource")),
    ("agent", ("agent", "user", "group", "service_account")),
    ("credential", ("credential",)),
    ("tool", ("tool",)),
    ("environment", ("environment",)),
    ("cluster", ("cluster",)),
)
ROLE_OF = {etype: role for role, etypes in _ROLE_GROUPS for etype in etypes}
_FINDING_TYPES = frozenset(_ROLE_GROUPS[0][1])
_UNKNOWN = "unknown"
# composite risk is on a 0-100 scale in new snapshots and 0-10 in old ones; the reference accepts either, so the lower bound decides
_RISK_BANDS = ((9.0, "critical"), (7.0, "high"), (4.0, "medium"))
_MCP_SEVERITY_ORDER = {"critical": 4, "high": 3, "medium": 2, "low": 1, "none": 0, "": 0}
# evidence block: wire key -> (node attribute, cast)
_EVIDENCE_FIELDS = (("cvssScore", "cvss_score", None), ("epssScore", "epss_score", None), ("isKev", "is_kev", bool), ("impactCategory", "impact_category", None))
_REST_ORIGIN = "graph_attack_path"
_MCP_ORIGIN = "mcp_exposure_paths"


def _etype(node) -> str:
    return enum_value(node.entity_type)


def _band(risk: float) -> str:
    for floor, name in _RISK_BANDS:
        if risk >= floor:
            return name
    return "none"


def _first_occurrences(items: Iterable[str]) -> list[str]:
    return list(dict.fromkeys(items))


def _stub(node_id: str) -> dict[str, Any]:
    return {"id": node_id, "label": node_id, "role": _UNKNOWN}


class EdgeIndex:
    """Hop-pair lookups over ``graph.edges`` (list order preserved), built once per graph instead of once per path."""

    __slots__ = ("edges", "_first", "_between", "_by_id")

    def __init__(self, edges):
        self.edges = list(edges or [])
        self._first: dict[tuple[str, str], int] = {}
        self._between: dict[tuple[str, str], list[int]] = {}
        self._by_id: dict[str, list[int]] = {}
        for pos, e in enumerate(self.edges):
            forward = (e.source, e.target)
            self._first.setdefault(forward, pos)
            if e.is_bidirectional:
                self._first.setdefault(forward[::-1], pos)
            self._between.setdefault((str(e.source), str(e.target)), []).append(pos)
            self._by_id.setdefault(str(getattr(e, "id", "")), []).append(pos)

    @classmethod
    def of(cls, edges) -> "EdgeIndex":
        return edges if isinstance(edges, cls) else cls(edges)

    def first_edge(self, source, target):
        pos = self._first.get((source, target))
        return None if pos is None else self.edges[pos]

    def positions_between(self, source, target) -> list[int]:
        return self._between.get((str(source), str(target)), [])

    def positions_with_id(self, edge_id) -> list[int]:
        return self._by_id.get(str(edge_id), [])


# ── REST (api/routes/graph.py) ────────────────────────────────────────────────────────────────────────────────────────
def exposure_role_for_node(node) -> str:
    return ROLE_OF.get(_etype(node), _UNKNOWN)


def exposure_ref_for_node(node_id: str, nodes_by_id: dict[str, Any]) -> dict[str, Any]:
    """One hop as the cockpit shows it; severity / riskScore only when the node carries them."""
    node = nodes_by_id.get(node_id)
    if node is None:
        return _stub(node_id)
    card: dict[str, Any] = {"id": node.id, "label": node.label, "role": exposure_role_for_node(node)}
    declared = getattr(node, "severity", "")
    if declared:
        card["severity"] = declared
    score = getattr(node, "risk_score", 0.0)
    if float(score or 0.0) > 0:
        card["riskScore"] = score
    return card


def finding_ids_for_nodes(nodes: dict[str, Any], path_hops: list[str], vuln_ids: list[str]) -> list[str]:
    """The path's own vulnerability ids (trimmed, blanks dropped), then the labels of finding-typed hops, first occurrence wins."""
    from_path = [v.strip() for v in vuln_ids]
    on_hops = []
    for hop in path_hops:
        node = nodes.get(hop)
        if node and _etype(node) in _FINDING_TYPES:
            on_hops.append(node.label or node.id)
    return _first_occurrences([v for v in from_path if v] + on_hops)


def _link(edge_id, source, target, relationship, direction, traversable, confidence) -> dict[str, Any]:
    return {"id": edge_id, "source": source, "target": target, "relationship": relationship, "direction": direction, "traversable": traversable,
            "confidence": confidence}


def exposure_relationships_for_path(path, edges) -> list[dict[str, Any]]:
    """One link per consecutive hop pair: the graph's first edge for the pair, else a synthetic directed link named after the path's own
    relationship at that position (``related`` past its end)."""
    index = EdgeIndex.of(edges)
    declared = list(path.edges)
    links = []
    for pos, (a, b) in enumerate(zip(path.hops, path.hops[1:])):
        hit = index.first_edge(a, b)
        if hit is None:
            name = declared[pos] if pos < len(declared) else "related"
            links.append(_link(f"{name}:{a}:{b}", a, b, name, "directed", True, 1.0))
        else:
            links.append(_link(hit.id, a, b, enum_value(hit.relationship), hit.direction, hit.traversable, getattr(hit, "confidence", 1.0)))
    return links


def _worst_declared_severity(hops, nodes_by_id: dict[str, Any], order: dict[str, int]) -> str:
    worst = ""
    for hop in hops:
        node = nodes_by_id.get(hop)
        declared = str(getattr(node, "severity", "") or "").lower() if node is not None else ""
        if order.get(declared, 0) > order.get(worst, 0):
            worst = declared
    return worst


def severity_for_exposure_path(path, nodes_by_id: dict[str, Any]) -> str:
    return _worst_declared_severity(path.hops, nodes_by_id, SEVERITY_RANK) or _band(path.composite_risk)


def _attributes(node) -> dict:
    return getattr(node, "attributes", {}) if node is not None else {}


def exposure_path_for_attack_path(path, *, nodes_by_id: dict[str, Any], edges=None, rank: int | None = None, scan_id: str = "") -> dict[str, Any]:
    """The ExposurePath object the cockpit and SDKs consume."""
    cards = [exposure_ref_for_node(h, nodes_by_id) for h in path.hops]
    blank = {"id": "", "label": "", "role": _UNKNOWN}
    head = exposure_ref_for_node(path.source, nodes_by_id) if path.source else (cards[0] if cards else blank)
    tail = exposure_ref_for_node(path.target, nodes_by_id) if path.target else (cards[-1] if cards else blank)
    links = exposure_relationships_for_path(path, edges)
    by_role: dict[str, list[dict]] = {}
    for card in cards:
        by_role.setdefault(card["role"], []).append(card)
    agents, servers, packages = by_role.get("agent", []), by_role.get("server", []), by_role.get("package", [])
    findings = finding_ids_for_nodes(nodes_by_id, path.hops, path.vuln_ids)
    what = findings[0] if findings else tail["label"]
    who = agents[0]["label"] if agents else head["label"]
    title = " via ".join(filter(None, (what, who))) or path.summary or "Exposure path"
    origin = {"source": _REST_ORIGIN}
    if scan_id:
        origin["scanId"] = scan_id
    fields = [
        ("id", "::".join((path.source, path.target, "->".join(path.hops)))), ("label", title), ("summary", path.summary), ("riskScore", path.composite_risk),
        ("severity", severity_for_exposure_path(path, nodes_by_id)), ("source", head), ("target", tail), ("hops", cards), ("relationships", links),
        ("nodeIds", list(path.hops)), ("edgeIds", [link["id"] for link in links]), ("findings", findings),
        ("affectedAgents", [c["label"] for c in agents]), ("affectedServers", [c["label"] for c in servers]),
        ("reachableTools", list(path.tool_exposure)), ("exposedCredentials", list(path.credential_exposure)), ("provenance", origin),
    ]
    if rank is not None:
        fields.append(("rank", rank))
    if packages or servers:
        pkg_node = nodes_by_id.get(packages[0]["id"]) if packages else None
        pkg_attrs = _attributes(pkg_node)
        fields.append(("dependencyContext", {
            "packageName": packages[0]["label"] if packages else "",
            "packageVersion": pkg_attrs.get("version", "") if pkg_node is not None else "",
            "ecosystem": pkg_attrs.get("ecosystem", "") if pkg_node is not None else "",
            "serverName": servers[0]["label"] if servers else "",
        }))
    finding = nodes_by_id.get(path.target)
    if finding is not None:
        attrs = _attributes(finding) or {}
        evidence = {key: (cast(attrs.get(attr)) if cast else attrs.get(attr)) for key, attr, cast in _EVIDENCE_FIELDS}
        evidence["source"] = _REST_ORIGIN
        fields.append(("evidence", evidence))
    return dict(fields)


def edge_relationships_for_hops(hops: list[str], edges) -> list[str]:
    """Relationship of the graph's first edge per consecutive hop pair; pairs without an edge contribute nothing."""
    index = EdgeIndex.of(edges)
    found = (index.first_edge(a, b) for a, b in zip(hops, hops[1:]))
    return [enum_value(e.relationship) for e in found if e is not None]


def serialize_attack_path(path, edges=None, *, nodes_by_id: dict[str, Any] | None = None, rank: int | None = None, scan_id: str = "") -> dict:
    """One element of the REST ``attack_paths`` list: the path's own dict, relationships filled from the graph when the path has none, and
    (given the node table) its ``exposure_path`` envelope."""
    record = path.to_dict()
    record["edges"] = [enum_value(e) for e in record.get("edges", [])]
    if edges is not None and not record["edges"]:
        record["edges"] = edge_relationships_for_hops(path.hops, edges)
    if nodes_by_id is not None:
        record["exposure_path"] = exposure_path_for_attack_path(path, nodes_by_id=nodes_by_id, edges=edges, rank=rank, scan_id=scan_id)
    return record


def serialize_attack_paths(graph, paths, *, first_rank: int = 1) -> list[dict]:
    """A page of ranked paths as the REST route emits it (ranks are 1-based positions in the ranked list); one edge index for all."""
    index = EdgeIndex(graph.edges)
    return [serialize_attack_path(p, index, nodes_by_id=graph.nodes, rank=first_rank + i, scan_id=graph.scan_id) for i, p in enumerate(paths)]


# ── MCP (mcp_tools/graph.py) ──────────────────────────────────────────────────────────────────────────────────────────
def _mcp_card(node_id: str, nodes_by_id: dict[str, Any]) -> dict[str, Any]:
    node = nodes_by_id.get(node_id)
    if node is None:
        return _stub(node_id)
    return {"id": node.id, "label": node.label, "role": _etype(node) or _UNKNOWN, "severity": getattr(node, "severity", ""),
            "riskScore": float(getattr(node, "risk_score", 0.0) or 0.0)}


def _mcp_links(path, edges) -> list[dict[str, Any]]:
    """Every edge named in ``path.edges`` by id or joining two consecutive hops, in ``graph.edges`` order, each once."""
    index = EdgeIndex.of(edges)
    hops = list(getattr(path, "hops", []) or [])
    chosen: set[int] = set()
    for edge_id in set(getattr(path, "edges", []) or []):
        chosen.update(index.positions_with_id(edge_id))
    for a, b in set(zip(hops, hops[1:])):
        chosen.update(index.positions_between(a, b))
    out = []
    for pos in sorted(chosen):
        e = index.edges[pos]
        out.append({"id": str(getattr(e, "id", "")), "source": str(e.source), "target": str(e.target), "relationship": enum_value(e.relationship),
                    "confidence": float(getattr(e, "confidence", 1.0) or 0.0)})
    return out


def mcp_exposure_path_payload(path, *, nodes_by_id: dict[str, Any], edges, rank: int, scan_id: str) -> dict[str, Any]:
    """The ``exposure_paths`` MCP tool's object for one ranked path (duck-typed: every path field is optional)."""
    def field(name, default):
        return getattr(path, name, default) or default

    hop_ids = list(field("hops", []))
    cards = [_mcp_card(h, nodes_by_id) for h in hop_ids]
    src, dst = field("source", ""), field("target", "")
    head = _mcp_card(str(src), nodes_by_id) if src else (cards[0] if cards else {})
    tail = _mcp_card(str(dst), nodes_by_id) if dst else (cards[-1] if cards else {})
    links = _mcp_links(path, edges)
    risk = float(field("composite_risk", 0.0))
    summary = getattr(path, "summary", "")
    return dict([
        ("id", "::".join((head.get("id", ""), tail.get("id", ""), "->".join(hop_ids)))), ("rank", rank), ("label", summary or "Exposure path"), ("summary", summary),
        ("riskScore", risk), ("severity", _worst_declared_severity(hop_ids, nodes_by_id, _MCP_SEVERITY_ORDER) or _band(risk)),
        ("source", head), ("target", tail), ("hops", cards), ("relationships", links), ("nodeIds", hop_ids),
        ("edgeIds", [link["id"] for link in links if link.get("id")]), ("findings", list(field("vuln_ids", []))),
        ("reachableTools", list(field("tool_exposure", []))), ("exposedCredentials", list(field("credential_exposure", []))),
        ("provenance", {"source": _MCP_ORIGIN, "scanId": scan_id}),
    ])

"""ExposurePath wire envelopes for attack paths — the step right after the path.

Restates, for ``AttackPath`` records produced on the device (``graph/exposure.py``) or loaded from a snapshot:

* the REST envelope ``_exposure_path_for_attack_path`` / ``_serialize_attack_path``
  (``/root/reference/src/agent_bom/api/routes/graph.py:506-670``, helpers ``:275-291``, ``:484-503``), and
* the MCP payload ``_exposure_path_payload`` (``/root/reference/src/agent_bom/mcp_tools/graph.py:16-100``).

The reference rebuilds an O(|E|) ``(source, target) → edge`` map for EVERY path (``routes/graph.py:543-547``,
``mcp_tools/graph.py:34-57``); here the two indexes are built once per graph (``EdgeIndex``) and give the same
answers: first edge in ``graph.edges`` order per pair (bidirectional edges also register the reverse pair) for REST,
every edge between consecutive hops in ``graph.edges`` order for MCP.  Pinned by ``tests/test_envelope.py`` on
envelopes the unmodified reference produced (``tests/golden/envelope/``, ``oracle/make_golden.py --envelope-only``).
"""

from __future__ import annotations

from typing import Any

from .schema import SEVERITY_RANK, enum_value

_FINDING = {"vulnerability", "misconfiguration"}
_ROLE = {
    "vulnerability": "finding", "misconfiguration": "finding", "package": "package",
    "server": "server", "container": "server", "cloud_resource": "server",
    "agent": "agent", "user": "agent", "group": "agent", "service_account": "agent",
    "credential": "credential", "tool": "tool", "environment": "environment", "cluster": "cluster",
}


def _type_value(node) -> str:
    return enum_value(node.entity_type)


def exposure_role_for_node(node) -> str:
    """routes/graph.py:506-524."""
    return _ROLE.get(_type_value(node), "unknown")


def exposure_ref_for_node(node_id: str, nodes_by_id: dict[str, Any]) -> dict[str, Any]:
    """routes/graph.py:527-540."""
    node = nodes_by_id.get(node_id)
    if node is None:
        return {"id": node_id, "label": node_id, "role": "unknown"}
    ref: dict[str, Any] = {"id": node.id, "label": node.label, "role": exposure_role_for_node(node)}
    if getattr(node, "severity", ""):
        ref["severity"] = node.severity
    if float(getattr(node, "risk_score", 0.0) or 0.0) > 0:
        ref["riskScore"] = node.risk_score
    return ref


def finding_ids_for_nodes(nodes: dict[str, Any], path_hops: list[str], vuln_ids: list[str]) -> list[str]:
    """routes/graph.py:275-291."""
    ids: list[str] = []
    seen: set[str] = set()
    for value in vuln_ids:
        cleaned = value.strip()
        if cleaned and cleaned not in seen:
            ids.append(cleaned)
            seen.add(cleaned)
    for hop in path_hops:
        node = nodes.get(hop)
        if not node or _type_value(node) not in _FINDING:
            continue
        label = node.label or node.id
        if label not in seen:
            ids.append(label)
            seen.add(label)
    return ids


class EdgeIndex:
    """``(source, target)`` lookups over ``graph.edges``, built once per graph instead of once per path."""

    def __init__(self, edges):
        self.edges = list(edges or [])
        self.first: dict[tuple[str, str], Any] = {}                  # REST: by_pair.setdefault (bidirectional edges also the reverse pair)
        self.all: dict[tuple[str, str], list[int]] = {}              # MCP: every edge of a directed pair, graph.edges order
        self.by_id: dict[str, list[int]] = {}
        for i, edge in enumerate(self.edges):
            self.first.setdefault((edge.source, edge.target), edge)
            if edge.is_bidirectional:
                self.first.setdefault((edge.target, edge.source), edge)
            self.all.setdefault((str(edge.source), str(edge.target)), []).append(i)
            self.by_id.setdefault(str(getattr(edge, "id", "")), []).append(i)


def _as_index(edges) -> EdgeIndex:
    return edges if isinstance(edges, EdgeIndex) else EdgeIndex(edges)


def exposure_relationships_for_path(path, edges) -> list[dict[str, Any]]:
    """routes/graph.py:543-580."""
    index = _as_index(edges)
    relationships: list[dict[str, Any]] = []
    for i, (source, target) in enumerate(zip(path.hops, path.hops[1:])):
        edge = index.first.get((source, target))
        if edge is not None:
            relationship = enum_value(edge.relationship)
            edge_id, direction, traversable, confidence = edge.id, edge.direction, edge.traversable, getattr(edge, "confidence", 1.0)
        else:
            relationship = path.edges[i] if i < len(path.edges) else "related"
            edge_id, direction, traversable, confidence = f"{relationship}:{source}:{target}", "directed", True, 1.0
        relationships.append({"id": edge_id, "source": source, "target": target, "relationship": relationship, "direction": direction,
                              "traversable": traversable, "confidence": confidence})
    return relationships


def _severity_from_risk(risk: float) -> str:
    if risk >= 90 or risk >= 9:
        return "critical"
    if risk >= 70 or risk >= 7:
        return "high"
    if risk >= 40 or risk >= 4:
        return "medium"
    return "none"


def severity_for_exposure_path(path, nodes_by_id: dict[str, Any]) -> str:
    """routes/graph.py:583-597."""
    severity = ""
    for hop in path.hops:
        node = nodes_by_id.get(hop)
        if node is not None and SEVERITY_RANK.get(str(getattr(node, "severity", "") or "").lower(), 0) > SEVERITY_RANK.get(severity, 0):
            severity = str(node.severity).lower()
    return severity or _severity_from_risk(path.composite_risk)


def exposure_path_for_attack_path(path, *, nodes_by_id: dict[str, Any], edges=None, rank: int | None = None, scan_id: str = "") -> dict[str, Any]:
    """routes/graph.py:597-656 — the ExposurePath object the cockpit and SDKs consume."""
    hops = [exposure_ref_for_node(hop, nodes_by_id) for hop in path.hops]
    empty_ref = {"id": "", "label": "", "role": "unknown"}
    source = exposure_ref_for_node(path.source, nodes_by_id) if path.source else (hops[0] if hops else empty_ref)
    target = exposure_ref_for_node(path.target, nodes_by_id) if path.target else (hops[-1] if hops else empty_ref)
    relationships = exposure_relationships_for_path(path, edges)
    packages = [hop for hop in hops if hop["role"] == "package"]
    servers = [hop for hop in hops if hop["role"] == "server"]
    agents = [hop for hop in hops if hop["role"] == "agent"]
    findings = finding_ids_for_nodes(nodes_by_id, path.hops, path.vuln_ids)
    label_parts = [findings[0] if findings else target["label"], agents[0]["label"] if agents else source["label"]]
    exposure: dict[str, Any] = {
        "id": f"{path.source}::{path.target}::{'->'.join(path.hops)}",
        "label": " via ".join(part for part in label_parts if part) or path.summary or "Exposure path",
        "summary": path.summary,
        "riskScore": path.composite_risk,
        "severity": severity_for_exposure_path(path, nodes_by_id),
        "source": source,
        "target": target,
        "hops": hops,
        "relationships": relationships,
        "nodeIds": list(path.hops),
        "edgeIds": [relationship["id"] for relationship in relationships],
        "findings": findings,
        "affectedAgents": [hop["label"] for hop in agents],
        "affectedServers": [hop["label"] for hop in servers],
        "reachableTools": list(path.tool_exposure),
        "exposedCredentials": list(path.credential_exposure),
        "provenance": {"source": "graph_attack_path", "scanId": scan_id} if scan_id else {"source": "graph_attack_path"},
    }
    if rank is not None:
        exposure["rank"] = rank
    if packages or servers:
        package_node = nodes_by_id.get(packages[0]["id"]) if packages else None
        exposure["dependencyContext"] = {
            "packageName": packages[0]["label"] if packages else "",
            "packageVersion": getattr(package_node, "attributes", {}).get("version", "") if package_node is not None else "",
            "ecosystem": getattr(package_node, "attributes", {}).get("ecosystem", "") if package_node is not None else "",
            "serverName": servers[0]["label"] if servers else "",
        }
    finding_node = nodes_by_id.get(path.target)
    if finding_node is not None:
        attributes = getattr(finding_node, "attributes", {}) or {}
        exposure["evidence"] = {
            "cvssScore": attributes.get("cvss_score"),
            "epssScore": attributes.get("epss_score"),
            "isKev": bool(attributes.get("is_kev")),
            "impactCategory": attributes.get("impact_category"),
            "source": "graph_attack_path",
        }
    return exposure


def edge_relationships_for_hops(hops: list[str], edges) -> list[str]:
    """routes/graph.py:488-503 — first relationship per consecutive hop pair; pairs without an edge are skipped."""
    if len(hops) < 2:
        return []
    index = _as_index(edges)
    out = []
    for source, target in zip(hops, hops[1:]):
        edge = index.first.get((source, target))
        if edge is not None:
            out.append(enum_value(edge.relationship))
    return out


def serialize_attack_path(path, edges=None, *, nodes_by_id: dict[str, Any] | None = None, rank: int | None = None, scan_id: str = "") -> dict:
    """routes/graph.py:659-670 — one element of the REST ``attack_paths`` list."""
    data = path.to_dict()
    data["edges"] = [enum_value(e) for e in data.get("edges", [])]
    if not data.get("edges") and edges is not None:
        data["edges"] = edge_relationships_for_hops(path.hops, edges)
    if nodes_by_id is not None:
        data["exposure_path"] = exposure_path_for_attack_path(path, nodes_by_id=nodes_by_id, edges=edges, rank=rank, scan_id=scan_id)
    return data


def serialize_attack_paths(graph, paths, *, first_rank: int = 1) -> list[dict]:
    """A page of ranked paths as the REST route emits it (ranks are 1-based positions in the ranked list); one edge index for all."""
    index = EdgeIndex(graph.edges)
    return [serialize_attack_path(p, index, nodes_by_id=graph.nodes, rank=first_rank + i, scan_id=graph.scan_id) for i, p in enumerate(paths)]


# ── MCP payload (mcp_tools/graph.py) ────────────────────────────────────────────────────────────────────────────────────
def _mcp_node_ref(node_id: str, nodes_by_id: dict[str, Any]) -> dict[str, Any]:
    node = nodes_by_id.get(node_id)
    if node is None:
        return {"id": node_id, "label": node_id, "role": "unknown"}
    return {"id": node.id, "label": node.label, "role": _type_value(node) or "unknown", "severity": getattr(node, "severity", ""),
            "riskScore": float(getattr(node, "risk_score", 0.0) or 0.0)}


def _mcp_relationship_refs(path, edges) -> list[dict[str, Any]]:
    """mcp_tools/graph.py:34-57 — every edge whose id is listed in ``path.edges`` or that joins two consecutive hops, graph.edges order."""
    index = _as_index(edges)
    hops = list(getattr(path, "hops", []) or [])
    picked: set[int] = set()
    for eid in set(getattr(path, "edges", []) or []):
        picked.update(index.by_id.get(str(eid), ()))
    for pair in set(zip(hops[:-1], hops[1:])):
        picked.update(index.all.get((str(pair[0]), str(pair[1])), ()))
    refs = []
    for i in sorted(picked):
        edge = index.edges[i]
        refs.append({"id": str(getattr(edge, "id", "")), "source": str(edge.source), "target": str(edge.target), "relationship": enum_value(edge.relationship),
                     "confidence": float(getattr(edge, "confidence", 1.0) or 0.0)})
    return refs


def _mcp_severity_for_path(path, nodes_by_id: dict[str, Any]) -> str:
    order = {"critical": 4, "high": 3, "medium": 2, "low": 1, "none": 0, "": 0}
    severity = ""
    for hop in getattr(path, "hops", []) or []:
        node = nodes_by_id.get(hop)
        candidate = str(getattr(node, "severity", "") or "").lower() if node is not None else ""
        if order.get(candidate, 0) > order.get(severity, 0):
            severity = candidate
    return severity or _severity_from_risk(float(getattr(path, "composite_risk", 0.0) or 0.0))


def mcp_exposure_path_payload(path, *, nodes_by_id: dict[str, Any], edges, rank: int, scan_id: str) -> dict[str, Any]:
    """mcp_tools/graph.py:78-100."""
    hops = [_mcp_node_ref(hop, nodes_by_id) for hop in getattr(path, "hops", []) or []]
    source = _mcp_node_ref(str(getattr(path, "source", "") or ""), nodes_by_id) if getattr(path, "source", "") else (hops[0] if hops else {})
    target = _mcp_node_ref(str(getattr(path, "target", "") or ""), nodes_by_id) if getattr(path, "target", "") else (hops[-1] if hops else {})
    relationships = _mcp_relationship_refs(path, edges)
    return {
        "id": f"{source.get('id', '')}::{target.get('id', '')}::{'->'.join(getattr(path, 'hops', []) or [])}",
        "rank": rank,
        "label": getattr(path, "summary", "") or "Exposure path",
        "summary": getattr(path, "summary", ""),
        "riskScore": float(getattr(path, "composite_risk", 0.0) or 0.0),
        "severity": _mcp_severity_for_path(path, nodes_by_id),
        "source": source,
        "target": target,
        "hops": hops,
        "relationships": relationships,
        "nodeIds": list(getattr(path, "hops", []) or []),
        "edgeIds": [relationship["id"] for relationship in relationships if relationship.get("id")],
        "findings": list(getattr(path, "vuln_ids", []) or []),
        "reachableTools": list(getattr(path, "tool_exposure", []) or []),
        "exposedCredentials": list(getattr(path, "credential_exposure", []) or []),
        "provenance": {"source": "mcp_exposure_paths", "scanId": scan_id},
    }

"""Inventory report → typed exposure graph (the load path in front of the traversal hot path).

Same entry point as the reference — ``build_unified_graph_from_report(report_json, *, scan_id, tenant_id)``
(``/root/reference/src/agent_bom/graph/builder.py:41-46``) — for the part of the report the exposure path is made
of: agents → MCP servers → packages / tools / credentials, package-level and blast-radius vulnerabilities, the
capability-impact and lateral-movement edges, model-provenance and dataset-card nodes.  Node ids, node order, edge
order, relationships, directions and first-wins de-duplication follow the reference section by section
(citations inline), so the CSR built from this graph is the CSR built from the reference's graph
(``tests/test_builder_identity.py`` checks ids / types / edges against the reference builder's output).

Deliberately light: nodes carry id / type / label / severity / risk score and a few attributes, not the
reference's provenance and compliance plumbing (≈35 % + 27 % of its build time).  Sections that
only the reference models (serving configs, CIS / SAST / IaC / skill-audit misconfigurations, framework topology,
cross-environment correlation, runtime session graph, toxic combinations) are NOT built here; their presence is
reported in ``graph.unhandled_sections`` so a caller can fall back to the reference builder +
``UnifiedGraph.from_graph`` when it needs them.

Two rules of the reference depend on keyword tables that live outside its graph package; they are injectable:
``classify_tool(name, description) -> list[str]`` (capabilities of a tool without declared ones; default: none)
and ``is_credential_key(name) -> bool`` (default: the reference builder's own fallback, builder.py:29-35).
"""

from __future__ import annotations

import re
from collections import defaultdict
from typing import Any, Callable, Iterable

from .container import UnifiedGraph
from .model import UnifiedEdge, UnifiedNode
from .schema import EntityType, RelationshipType

UNHANDLED_SECTIONS = (
    "serving_configs", "cis_benchmark_data", "snowflake_cis_benchmark_data", "azure_cis_benchmark_data", "gcp_cis_benchmark_data", "aws_cis_benchmark",
    "sast_data", "iac_findings_data", "skill_audit", "ai_inventory", "runtime_session_graph", "toxic_combinations",
)
SEVERITY_RISK_SCORE = {"critical": 8.0, "high": 6.0, "medium": 4.0, "low": 2.0}   # edge weights only; not used by traversal
_PYPI_SEP = re.compile(r"[-_.]+")
_CAPABILITY_ALIASES = {"readonly": "read", "read_only": "read", "destructive": "delete", "exec": "execute", "execution": "execute",
                       "network_egress": "network", "egress": "network", "credential": "auth", "credentials": "auth", "administrative": "admin"}
_CAPABILITIES = {"read", "write", "delete", "execute", "network", "auth", "admin"}
_UNMAPPABLE_VERSIONS = {"unknown", "latest", "*", "main", "master"}


def _default_is_credential_key(name: str) -> bool:
    low = name.lower()
    return any(p in low for p in ("key", "token", "secret", "password", "auth"))


# ── package identity (reference package_utils.py:39-117, version_utils.py:82-107) ──

def normalize_ecosystem(ecosystem: str) -> str:
    eco = (ecosystem or "").strip().lower()
    return {"golang": "go"}.get(eco, eco)


def normalize_name(name: str, ecosystem: str) -> str:
    if not name:
        return name
    return _PYPI_SEP.sub("-", name).lower() if ecosystem == "pypi" else name.lower()


def normalize_version(version: str, ecosystem: str) -> str:
    version = (version or "").strip()
    if not version or version in ("latest", "unknown"):
        return version
    if ecosystem != "go" and version.startswith("v"):
        version = version[1:]
    if ecosystem == "pypi":
        version = re.sub(r"\.?(alpha|a)(\d+)?", r"a\2", version, flags=re.IGNORECASE)
        version = re.sub(r"\.?(beta|b)(\d+)?", r"b\2", version, flags=re.IGNORECASE)
        version = re.sub(r"\.?(preview|c|rc)(\d+)?", r"rc\2", version, flags=re.IGNORECASE)
        version = re.sub(r"\.?(post|rev|r)(\d+)?", r".post\2", version, flags=re.IGNORECASE)
        version = re.sub(r"\.?(dev)(\d+)?", r".dev\2", version, flags=re.IGNORECASE)
    return version


def package_key(name: str, version: str, ecosystem: str, purl: str | None = None) -> str:
    """``canonical_package_key`` without purl parsing (a report that carries purls needs the reference builder)."""
    if purl:
        raise NotImplementedError("package identities taken from a purl are resolved by the reference builder only")
    eco = normalize_ecosystem(ecosystem)
    nm = normalize_name((name or "").strip(), eco)
    ver = normalize_version(version, eco)
    return f"{eco}:{nm}@{ver}" if ver else f"{eco}:{nm}"


def _tool_capabilities(tool: dict[str, Any], classify_tool: Callable[[str, str], Iterable[str]] | None) -> list[str]:
    """Declared capabilities win; otherwise the injected classifier + schema findings (reference builder.py:900-925)."""
    declared_values = tool.get("capabilities") or tool.get("declared_capabilities") or []
    if isinstance(declared_values, list):
        declared = set()
        for raw in declared_values:
            if isinstance(raw, str):
                norm = raw.strip().lower().replace("-", "_").replace(" ", "_")
                norm = _CAPABILITY_ALIASES.get(norm, norm)
                if norm in _CAPABILITIES:
                    declared.add(norm)
        if declared:
            return sorted(declared)
    caps = set(classify_tool(str(tool.get("name", "")), str(tool.get("description", ""))) if classify_tool else ())
    findings = tool.get("schema_findings", [])
    if isinstance(findings, list):
        for finding in findings:
            low = str(finding).lower()
            if "network-egress" in low or "url" in low:
                caps.add("network")
            if "shell-execution" in low or "command" in low:
                caps.add("execute")
            if "filesystem" in low or "path" in low:
                caps.add("read")
    return sorted(caps)


def _mappable(version: Any) -> bool:
    v = str(version or "").strip().lower()
    return bool(v and v not in _UNMAPPABLE_VERSIONS)


def build_unified_graph_from_report(report_json: dict[str, Any], *, scan_id: str = "", tenant_id: str = "", device: int = 0,
                                    classify_tool: Callable[[str, str], Iterable[str]] | None = None,
                                    is_credential_key: Callable[[str], bool] | None = None, columnar: bool = False) -> UnifiedGraph:
    """``columnar=True`` builds the graph as COLUMNS (``graph/columnar.py``): no ``UnifiedNode`` / ``UnifiedEdge`` object is made
    per node or edge, the CSR comes straight from the index arrays, and records are synthesised only for what a caller touches."""
    is_cred = is_credential_key or _default_is_credential_key
    if columnar:
        from .columnar import ColumnSink

        sink = ColumnSink(scan_id=scan_id or report_json.get("scan_id", ""), tenant_id=tenant_id, device=device)
    else:
        sink = _ObjectSink(UnifiedGraph(scan_id=scan_id or report_json.get("scan_id", ""), tenant_id=tenant_id, device=device))
    unhandled = [k for k in UNHANDLED_SECTIONS if report_json.get(k)]
    agents_data = report_json.get("agents", [])
    blast_data = report_json.get("blast_radius", report_json.get("blast_radii", []))
    scan_sources = report_json.get("scan_sources", [])
    source_tag = scan_sources[0] if scan_sources else "mcp-scan"

    server_to_agents: dict[str, list[str]] = defaultdict(list)
    cred_to_agents: dict[str, list[str]] = defaultdict(list)
    pkg_key_to_servers: dict[str, list[str]] = defaultdict(list)
    server_name_to_agent_servers: dict[str, dict[str, str]] = defaultdict(dict)
    agent_to_server_ids: dict[str, set[str]] = defaultdict(set)
    server_to_tool_ids: dict[str, list[str]] = defaultdict(list)
    package_id_to_servers: dict[str, list[str]] = defaultdict(list)
    tool_has_caps: dict[str, bool] = {}
    pending: list[tuple[str, str, Any, str]] = []     # (vuln node, server, package version, severity) for capability-impact edges

    def node(nid: str, et: EntityType, label: str, **kw) -> None:
        sink.node(nid, et, label, source_tag, **kw)

    edge = sink.edge

    def exploitable_via(vuln_id: str, server_id: str, version: Any, severity: str) -> None:
        """vuln -> every tool of the server that has capabilities, if the package version is mappable (builder.py:971-1019)."""
        if not _mappable(version):
            return
        for tool_id in server_to_tool_ids.get(server_id, []):
            if tool_has_caps.get(tool_id):
                edge(vuln_id, tool_id, RelationshipType.EXPLOITABLE_VIA, weight=SEVERITY_RISK_SCORE.get(severity, 1.0))

    # ── agents → servers → packages → tools → credentials (builder.py:81-370) ──
    for agent in agents_data:
        agent_name = agent.get("name", "unknown")
        scope = ""
        for key in ("source_id", "endpoint_id", "device_id"):                      # builder.py:1818-1830
            scope = str(agent.get(key) or "").strip()
            if scope:
                break
        meta = agent.get("metadata") if isinstance(agent.get("metadata"), dict) else {}
        if not scope:
            for key in ("source_id", "endpoint_id", "device_id"):
                scope = str(meta.get(key) or "").strip()
                if scope:
                    break
        name_part = str(agent_name or "unknown").strip() or "unknown"
        agent_id = f"agent:{scope.replace(':', '%3A')}:{name_part}" if scope else f"agent:{name_part}"      # builder.py:1833-1839
        agent_key = agent_id.removeprefix("agent:")
        provider_name = str(agent.get("source") or "local").strip() or "local"
        provider_id = f"provider:{provider_name}"
        if isinstance(meta.get("cloud_origin"), dict) and "cloud_lineage" not in unhandled:
            unhandled.append("cloud_lineage")
        node(provider_id, EntityType.PROVIDER, provider_name)
        node(agent_id, EntityType.AGENT, agent_name, attributes={"agent_type": agent.get("type", agent.get("agent_type", ""))})
        edge(provider_id, agent_id, RelationshipType.HOSTS)

        for srv in agent.get("mcp_servers", []):
            srv_name = srv.get("name", "unknown")
            srv_id = f"server:{agent_key}:{srv_name}"
            node(srv_id, EntityType.SERVER, srv_name, attributes={"agent": agent_name})
            edge(agent_id, srv_id, RelationshipType.USES)
            server_to_agents[srv_name].append(agent_id)
            server_name_to_agent_servers[srv_name][agent_id] = srv_id
            agent_to_server_ids[agent_name].add(srv_id)
            if scope:
                agent_to_server_ids[scope].add(srv_id)
                agent_to_server_ids[f"{scope}:{agent_name}"].add(srv_id)

            for pkg in srv.get("packages", []):
                pkg_name, pkg_version, eco = pkg.get("name", "unknown"), pkg.get("version", ""), pkg.get("ecosystem", "")
                key = package_key(str(pkg_name or "unknown"), str(pkg_version or ""), str(eco or ""), pkg.get("purl"))
                pkg_id = f"pkg:{key}"
                node(pkg_id, EntityType.PACKAGE, f"{pkg_name}@{pkg_version}" if pkg_version else pkg_name, attributes={"version": pkg_version, "ecosystem": eco})
                edge(srv_id, pkg_id, RelationshipType.DEPENDS_ON)
                package_id_to_servers[pkg_id].append(srv_id)
                pkg_key_to_servers[package_key(pkg_name, pkg_version, eco, pkg.get("purl"))].append(srv_id)
                for vuln in pkg.get("vulnerabilities", []):                          # builder.py:271-283, 1036-1072
                    vid = vuln.get("id", "")
                    if not vid:
                        continue
                    severity = vuln.get("severity", "").lower()
                    vuln_id = f"vuln:{vid}"
                    node(vuln_id, EntityType.VULNERABILITY, vid, severity=severity)
                    edge(pkg_id, vuln_id, RelationshipType.VULNERABLE_TO, weight=SEVERITY_RISK_SCORE.get(severity, 1.0))
                    pending.append((vuln_id, srv_id, pkg.get("version", ""), str(vuln.get("severity", "") or "").lower()))

            tool_ids: list[str] = []
            for tool in srv.get("tools", []):
                tool_name = tool.get("name", "unknown")
                tool_id = f"tool:{srv_id}:{tool_name}"
                tool_ids.append(tool_id)
                caps = _tool_capabilities(tool, classify_tool)
                node(tool_id, EntityType.TOOL, tool_name, attributes={"capabilities": caps, "server": srv_id, "agent": agent_name})
                tool_has_caps[tool_id] = bool([c for c in caps if str(c)])     # add_node merges attributes: the stored list is this one
                server_to_tool_ids[srv_id].append(tool_id)
                edge(srv_id, tool_id, RelationshipType.PROVIDES_TOOL)

            env_keys = srv.get("credential_env_vars", [])
            if not env_keys:
                env = srv.get("env", {})
                if isinstance(env, dict):
                    env_keys = [k for k in env if is_cred(k)]
            for env_key in env_keys:
                cred_id = f"cred:{env_key}"
                node(cred_id, EntityType.CREDENTIAL, env_key)
                edge(srv_id, cred_id, RelationshipType.EXPOSES_CRED, weight=2.0)
                cred_to_agents[env_key].append(agent_id)
                for tool_id in tool_ids:
                    edge(cred_id, tool_id, RelationshipType.REACHES_TOOL)

    for vuln_id, srv_id, version, severity in pending:                              # builder.py:371-382
        exploitable_via(vuln_id, srv_id, version, severity)

    # ── blast-radius vulnerabilities (builder.py:384-471) ──
    for br in blast_data:
        vid = br.get("vulnerability_id", "")
        if not vid:
            continue
        severity = br.get("severity", "").lower()
        pkg_name = br.get("package_name", br.get("package", "").split("@")[0])
        pkg_version, eco = br.get("package_version", ""), br.get("ecosystem", "")
        purl = br.get("package_purl") or br.get("purl")
        vuln_id = f"vuln:{vid}"
        node(vuln_id, EntityType.VULNERABILITY, vid, severity=severity, risk_score=br.get("risk_score", 0))
        pkg_id = f"pkg:{package_key(pkg_name, pkg_version, eco, purl)}" if pkg_name else ""
        if pkg_name and sink.has_node(pkg_id):
            edge(pkg_id, vuln_id, RelationshipType.VULNERABLE_TO, weight=SEVERITY_RISK_SCORE.get(severity, 1.0))
        # affected servers: package hosts, narrowed by named servers, narrowed by named agents (builder.py:1083-1137)
        candidates: set[str] = set()
        if pkg_name:
            candidates.update(pkg_key_to_servers.get(package_key(pkg_name, pkg_version, eco, purl), []))
        names = set()
        for s in br.get("affected_servers", []):
            nm = str(s.get("name", "")).strip() if isinstance(s, dict) else str(getattr(s, "name", s)).strip()
            if nm:
                names.add(nm)
        if names:
            named: set[str] = set()
            for nm in names:
                named.update(server_name_to_agent_servers.get(nm, {}).values())
            candidates = (candidates & named) if candidates else named
        agent_names = {str(a).strip() for a in br.get("affected_agents", []) if str(a).strip()}
        if agent_names:
            by_agent: set[str] = set()
            for nm in agent_names:
                by_agent.update(agent_to_server_ids.get(nm, set()))
            candidates = (candidates & by_agent) if candidates else by_agent
        affected = sorted(candidates)
        for srv_id in affected:
            edge(srv_id, vuln_id, RelationshipType.VULNERABLE_TO, weight=SEVERITY_RISK_SCORE.get(severity, 1.0))
        for srv_id in affected:                                                     # builder.py:454-471, 948-968
            if pkg_name and srv_id in package_id_to_servers.get(pkg_id, []) and _mappable(br.get("package_version")):
                exploitable_via(vuln_id, srv_id, br.get("package_version"), severity)

    # ── lateral movement: shared servers / shared credentials, agent ↔ agent (builder.py:475-507) ──
    for rel, groups, weight in ((RelationshipType.SHARES_SERVER, server_to_agents, 3.0), (RelationshipType.SHARES_CRED, cred_to_agents, 4.0)):
        done: set[tuple[str, ...]] = set()
        for members in groups.values():
            unique = sorted(set(members))
            key = tuple(unique)
            if key in done:          # a second group with the very same members only repeats (source, target, relationship) triples: first wins
                continue
            done.add(key)
            for i, a1 in enumerate(unique):
                for a2 in unique[i + 1:]:
                    edge(a1, a2, rel, direction="bidirectional", weight=weight)

    # ── model provenance / dataset cards: nodes only (builder.py:509-554) ──
    for m in report_json.get("model_provenance", []):
        name = m.get("model_name", m.get("name", "unknown"))
        sink.node(f"model:{name}", EntityType.MODEL, name, "model-provenance")
    cards = report_json.get("dataset_cards")
    if isinstance(cards, dict):
        for d in cards.get("datasets", []):
            name = d.get("name") or d.get("source_file") or "unknown-dataset"
            sink.node(f"dataset:{name}", EntityType.DATASET, name, "dataset-cards")
    g = sink.finish()
    g.unhandled_sections = unhandled
    return g


class _ObjectSink:
    """Default sink: one ``UnifiedNode`` / ``UnifiedEdge`` record per node / edge, merged by ``UnifiedGraph.add_node`` / ``add_edge``."""

    def __init__(self, graph: UnifiedGraph):
        self.g = graph

    def node(self, nid: str, et: EntityType, label: str, source: str, **kw) -> None:
        self.g.add_node(UnifiedNode(id=nid, entity_type=et, label=label, data_sources=[source], **kw))

    def edge(self, src: str, dst: str, rel: RelationshipType, **kw) -> None:
        self.g.add_edge(UnifiedEdge(source=src, target=dst, relationship=rel, **kw))

    def has_node(self, nid: str) -> bool:
        return self.g.has_node(nid)

    def finish(self) -> UnifiedGraph:
        return self.g

"""The light inventory builder makes the graph the reference builder makes (ids, order, types, labels, severities, edges).

Fixture: tests/golden/identity/builder_identity.json.gz (oracle/make_golden.py runs the UNMODIFIED reference builder on
each report and records its nodes / edges, plus the reference's own tool-classification and credential-key answers for
the names in the report - the two keyword tables the reference keeps outside its graph package).  CPU-only.
"""

from __future__ import annotations

import gzip
import json
from pathlib import Path

import pytest

FIX = Path(__file__).resolve().parent / "golden" / "identity" / "builder_identity.json.gz"
DOCS = json.loads(gzip.open(FIX, "rb").read())
SEP = "|#|"


@pytest.mark.parametrize("doc", DOCS, ids=[d["label"] for d in DOCS])
def test_builder_matches_reference(doc):
    from agent_bom_b200.graph import build_unified_graph_from_report
    from agent_bom_b200.graph.schema import enum_value

    caps = doc["tool_caps"]
    creds = set(doc["cred_keys"])
    g = build_unified_graph_from_report(json.loads(json.dumps(doc["report"])), classify_tool=lambda n, d: caps.get(n + SEP + d, []),
                                        is_credential_key=lambda k: k in creds)
    assert g.unhandled_sections == []
    got_nodes = [[n.id, enum_value(n.entity_type), n.label, n.severity, float(n.risk_score or 0.0)] for n in g.nodes.values()]
    got_edges = [[e.source, e.target, enum_value(e.relationship), e.direction, bool(e.traversable)] for e in g.edges]
    assert got_nodes == doc["nodes"]
    assert got_edges == doc["edges"]


def test_unmodelled_sections_are_reported_not_guessed():
    from agent_bom_b200.graph import build_unified_graph_from_report

    g = build_unified_graph_from_report({"agents": [{"name": "a", "mcp_servers": []}], "runtime_session_graph": {"nodes": [1]}, "sast_data": {"findings": [1]}})
    assert set(g.unhandled_sections) == {"runtime_session_graph", "sast_data"}
    assert list(g.nodes) == ["provider:local", "agent:a"]
    with pytest.raises(NotImplementedError):
        build_unified_graph_from_report({"agents": [{"name": "a", "mcp_servers": [{"name": "s", "packages": [{"name": "p", "purl": "pkg:npm/p@1"}]}]}]})

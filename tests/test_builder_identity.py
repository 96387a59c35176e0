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


@pytest.mark.parametrize("columnar", [False, True], ids=["records", "columnar"])
@pytest.mark.parametrize("doc", DOCS, ids=[d["label"] for d in DOCS])
def test_builder_matches_reference(doc, columnar):
    from agent_bom_b200.graph import build_unified_graph_from_report
    from agent_bom_b200.graph.schema import enum_value

    caps = doc["tool_caps"]
    creds = set(doc["cred_keys"])
    g = build_unified_graph_from_report(json.loads(json.dumps(doc["report"])), classify_tool=lambda n, d: caps.get(n + SEP + d, []),
                                        is_credential_key=lambda k: k in creds, columnar=columnar)
    assert g.unhandled_sections == []
    got_nodes = [[n.id, enum_value(n.entity_type), n.label, n.severity, float(n.risk_score or 0.0)] for n in g.nodes.values()]
    got_edges = [[e.source, e.target, enum_value(e.relationship), e.direction, bool(e.traversable)] for e in g.edges]
    assert got_nodes == doc["nodes"]
    assert got_edges == doc["edges"]
    if columnar:      # the CSR built straight from the columns is the CSR of the record graph; no record was made to get it
        from agent_bom_b200.graph import csr as csrmod
        from agent_bom_b200.graph.columnar import ColumnarGraph

        g2 = build_unified_graph_from_report(json.loads(json.dumps(doc["report"])), classify_tool=lambda n, d: caps.get(n + SEP + d, []),
                                             is_credential_key=lambda k: k in creds, columnar=True)
        assert isinstance(g2, ColumnarGraph) and not g2.nodes._cache
        ref = csrmod.from_unified_graph(build_unified_graph_from_report(json.loads(json.dumps(doc["report"])), classify_tool=lambda n, d: caps.get(n + SEP + d, []),
                                                                         is_credential_key=lambda k: k in creds))
        for name in ("fwd_off", "fwd_nbr", "fwd_meta", "fwd_eid", "rev_off", "rev_nbr", "rev_meta", "rev_eid", "node_type", "node_rank"):
            assert (getattr(g2.csr, name) == getattr(ref, name)).all(), name
        assert g2.csr.node_ids == ref.node_ids and not g2.nodes._cache
        # touching one node synthesises exactly one record; mutation falls back to ordinary records
        first = next(iter(g2.nodes))
        assert g2.nodes[first].id == first and len(g2.nodes._cache) == 1
        from agent_bom_b200.graph import UnifiedNode

        g2.add_node(UnifiedNode(id="agent:new", entity_type="agent", label="new"))
        assert isinstance(g2.nodes, dict) and "agent:new" in g2.nodes and len(g2.edges) == len(doc["edges"])


def test_unmodelled_sections_are_reported_not_guessed():
    from agent_bom_b200.graph import build_unified_graph_from_report

    g = build_unified_graph_from_report({"agents": [{"name": "a", "mcp_servers": []}], "runtime_session_graph": {"nodes": [1]}, "sast_data": {"findings": [1]}})
    assert set(g.unhandled_sections) == {"runtime_session_graph", "sast_data"}
    assert list(g.nodes) == ["provider:local", "agent:a"]
    with pytest.raises(NotImplementedError):
        build_unified_graph_from_report({"agents": [{"name": "a", "mcp_servers": [{"name": "s", "packages": [{"name": "p", "purl": "pkg:npm/p@1"}]}]}]})

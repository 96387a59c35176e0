"""f2: ExposurePath wire envelopes (REST + MCP) equal the unmodified reference's, object for object.

Host-side Python — no GPU needed: the inputs are ``AttackPath`` records (here taken from the golden file, in the
device pipeline they come from ``graph/exposure.py``) plus the graph's node and edge records.
Goldens: ``oracle/make_golden.py --envelope-only`` → ``tests/golden/envelope/envelopes.json.gz``
(reference api/routes/graph.py:597-670, mcp_tools/graph.py:78-100)."""

from __future__ import annotations

import gzip
import json
from pathlib import Path

import pytest

from agent_bom_b200.graph import AttackPath, UnifiedEdge, UnifiedGraph, UnifiedNode
from agent_bom_b200.graph.envelope import EdgeIndex, mcp_exposure_path_payload, serialize_attack_path, serialize_attack_paths

DOC = json.loads(gzip.open(Path(__file__).resolve().parent / "golden" / "envelope" / "envelopes.json.gz", "rb").read())


def build(case) -> UnifiedGraph:
    g = UnifiedGraph(scan_id=case["scan_id"], tenant_id="default")
    for n in case["nodes"]:
        g.nodes[n["id"]] = UnifiedNode(id=n["id"], entity_type=n["entity_type"], label=n["label"], severity=n["severity"], risk_score=n["risk_score"], attributes=n["attributes"])
    for e in case["edges"]:
        g.edges.append(UnifiedEdge(source=e["source"], target=e["target"], relationship=e["relationship"], direction=e["direction"], traversable=e["traversable"]))
    return g


@pytest.mark.parametrize("case", DOC["cases"], ids=[c["name"] for c in DOC["cases"]])
def test_rest_and_mcp_envelopes_equal_the_reference(case):
    g = build(case)
    paths = [AttackPath(**p) for p in case["paths"]]
    index = EdgeIndex(g.edges)           # one index for the whole page (the reference rebuilds the pair map per path)
    for p, rank, want_rest, want_mcp in zip(paths, case["ranks"], case["rest"], case["mcp"]):
        assert serialize_attack_path(p, index, nodes_by_id=g.nodes, rank=rank, scan_id=g.scan_id) == want_rest
        assert serialize_attack_path(p, g.edges, nodes_by_id=g.nodes, rank=rank, scan_id=g.scan_id) == want_rest      # plain edge list works too
        assert mcp_exposure_path_payload(p, nodes_by_id=g.nodes, edges=index, rank=rank, scan_id=g.scan_id) == want_mcp
    for p, want in zip(paths, case["rest_without_nodes"]):
        assert serialize_attack_path(p, g.edges) == want
    if case["ranks"] == list(range(1, len(paths) + 1)):
        assert serialize_attack_paths(g, paths) == case["rest"]

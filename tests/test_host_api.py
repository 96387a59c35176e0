"""CPU-only tests of the host layer: container semantics, CSR export, path ranking, store conventions, sharding + gloo broadcast."""

from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from golden_util import ALL_FIXTURES, edge_arrays, graph_from_fixture, load, node_rank, oracle_graph
from oracle import oracle as orc

ROOT = Path(__file__).resolve().parents[1]


def test_add_node_and_add_edge_follow_the_reference_rules():
    """container.py:116-198 — merge rules and (source,target,relationship) de-duplication with evidence merge."""
    from agent_bom_b200.graph import EntityType, RelationshipType, UnifiedEdge, UnifiedGraph, UnifiedNode

    g = UnifiedGraph()
    g.add_node(UnifiedNode(id="a", entity_type=EntityType.AGENT, label="a", risk_score=1.0, severity="low", data_sources=["x"]))
    g.add_node(UnifiedNode(id="a", entity_type=EntityType.AGENT, label="a", risk_score=5.0, severity="critical", data_sources=["x", "y"], attributes={"k": 1}))
    assert len(g.nodes) == 1 and g.nodes["a"].risk_score == 5.0 and g.nodes["a"].severity == "critical"
    assert g.nodes["a"].data_sources == ["x", "y"] and g.nodes["a"].attributes == {"k": 1}
    e = UnifiedEdge(source="a", target="b", relationship=RelationshipType.USES, evidence={"cvss": None})
    g.add_edge(e)
    g.add_edge(UnifiedEdge(source="a", target="b", relationship=RelationshipType.USES, evidence={"cvss": 9.8, "empty": ""}))
    assert len(g.edges) == 1 and g.edges[0].evidence == {"cvss": 9.8}
    g.add_edge(UnifiedEdge(source="a", target="b", relationship=RelationshipType.DEPENDS_ON))
    assert len(g.edges) == 2


@pytest.mark.parametrize("name", ["kat_probe", "kat_schema", "mesh_inventory"])
def test_container_csr_export_matches_fixture(name):
    """Node order, ghost endpoints, codes and both CSRs from the container equal the fixture's index-space graph."""
    doc = load(name)
    g = graph_from_fixture(doc)
    c = g.csr
    assert c.node_ids == doc["node_ids"] and c.n_real == doc["n_real"]
    np.testing.assert_array_equal(c.node_type, np.asarray(doc["node_types"], dtype=np.uint8))
    src, dst, rel, flags = edge_arrays(doc)
    np.testing.assert_array_equal(c.src, src); np.testing.assert_array_equal(c.dst, dst)
    np.testing.assert_array_equal(c.rel, rel); np.testing.assert_array_equal(c.flags, flags)
    og = oracle_graph(name)
    for arr in ("fwd_off", "fwd_nbr", "fwd_meta", "fwd_eid", "rev_off", "rev_nbr", "rev_meta", "rev_eid"):
        np.testing.assert_array_equal(getattr(c, arr), getattr(og, arr), err_msg=arr)
    np.testing.assert_array_equal(c.node_rank, node_rank(doc["node_ids"]))
    # point queries come from the CSR rows, in list order
    if doc.get("adjacency"):
        ids = doc["node_ids"]
        for u, lst in doc["adjacency"].items():
            assert [e.target for e in g.edges_from(ids[int(u)])] == [ids[v] for v, _ in lst]
        for u, lst in doc["reverse_adjacency"].items():
            assert g.sources_of(ids[int(u)]) == [ids[v] for v, _ in lst]


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_attack_path_materialisation_and_ranking(name):
    """Host half of _derived_attack_paths (risk floats, labels, stable sort) on oracle-produced rows == reference output."""
    from agent_bom_b200.graph import materialize_attack_paths
    from agent_bom_b200.graph.schema import REL_CODE

    doc = load(name)
    g = graph_from_fixture(doc)
    rows = orc.derived_paths(oracle_graph(name), doc["findings"], node_rank(doc["node_ids"]))
    got = materialize_attack_paths(g, rows)
    want = doc["cases"]["derived_paths"]
    ids = doc["node_ids"]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.hops == [ids[h] for h in b["hops"]] and [REL_CODE[e] for e in a.edges] == b["edges"]
        assert a.composite_risk == b["risk"] and a.credential_exposure == b["creds"] and a.tool_exposure == b["tools"] and a.vuln_ids == b["vuln_ids"]
        assert a.source == ids[b["source"]] and a.target == ids[b["target"]]


def test_store_conventions_without_a_snapshot():
    """Missing snapshot / node -> None, ([], set()), empty graph; unknown ops raise the typed error (api/graph_store.py:644-646,703-704)."""
    from agent_bom_b200.store import B200GraphStore, B200UnsupportedOperationError

    s = B200GraphStore()
    assert s.impact_of(node_id="x") is None
    assert s.bfs_paths(source="x") == ([], set())
    sub, depth, trunc = s.traverse_subgraph(roots=["x"])
    assert not sub.nodes and depth == {} and trunc is False
    assert s.attack_paths() == ("", "", [], 0) and s.attack_paths_for_sources(source_ids={"a"}) == []
    with pytest.raises(B200UnsupportedOperationError):
        s.list_snapshots(tenant_id="t")


def test_shard_bounds_cover_everything_once():
    from agent_bom_b200.dist import shard_bounds

    for n in (0, 1, 7, 8, 9, 1000, 3_740_035):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("name", ["kat_chain", "kat_reach_misc", "mesh_inventory", "estate_dense_40"])
def test_dependency_reach_partials_merge_to_the_whole(name):
    """a9 across ranks: agent shards of any shape merge (union by id rank, minimum over reaching shards only) to the unsplit answer."""
    from agent_bom_b200.dist import REACH_KEYS, merge_dependency_reach
    from agent_bom_b200.graph.schema import REACH_MASK, VULN_PKG_MASK

    doc = load(name)
    og = oracle_graph(name)
    rank = node_rank(doc["node_ids"])
    agents = np.flatnonzero(og.node_type == 0).astype(np.int32)
    whole = orc.dependency_reach(og, agents, REACH_MASK, VULN_PKG_MASK, rank)
    for cuts in ([0, len(agents)], [0, 1, len(agents)], [0, len(agents) // 3, len(agents) // 3, len(agents)], list(range(len(agents) + 1))[:6] + [len(agents)]):
        parts = [orc.dependency_reach(og, agents[a:b], REACH_MASK, VULN_PKG_MASK, rank) for a, b in zip(cuts, cuts[1:])]
        merged = merge_dependency_reach(parts, rank)
        for key in REACH_KEYS:
            assert np.array_equal(np.asarray(merged[key]), np.asarray(whole[key])), (key, cuts)


def test_gloo_world2_broadcast_and_shard():
    """N>1 plumbing on CPU: rank 0's CSR reaches rank 1 bit-identically over a world_size-2 gloo group; shards tile the source list."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29533",
                           str(ROOT / "tests" / "_gloo_worker.py")], capture_output=True, text=True, env=env, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "rank0-ok" in proc.stdout and "rank1-ok" in proc.stdout
    # dependency reach split over the two ranks (per-package union / minimum exchanged by all-gather) equals the unsplit answer
    assert "rank0-reach-ok" in proc.stdout and "rank1-reach-ok" in proc.stdout


def test_walk_result_rebuilds_the_dense_histogram_from_packed_columns():
    """The host result carries only the entity-type columns present in the batch (csrc/histpack.cuh); `hist` rebuilds the dense
    table on first access and `hist_dict` (the reference's `affected_by_type`) reads the packed rows directly."""
    from agent_bom_b200.engine import WalkResult
    from agent_bom_b200.graph.schema import ENTITY_VALUES, N_ENTITY_TYPES

    rng = np.random.default_rng(5)
    dense = np.zeros((50, N_ENTITY_TYPES), np.uint32)
    cols = np.array([0, 2, 3, 9, 17], dtype=np.int64)
    dense[:, cols] = rng.integers(0, 400, size=(50, len(cols)))
    dense[7] = 0
    z = np.zeros(50, np.int32)
    for dt in (np.uint16, np.uint32):
        r = WalkResult(start=z.astype(np.int64), count=z, maxd=z, flags=z, nodes=np.zeros(0, np.int32), hist_packed=dense[:, cols].astype(dt), hist_cols=cols)
        assert r._hist is None
        assert r.hist_dict(3) == {ENTITY_VALUES[t]: int(c) for t, c in enumerate(dense[3]) if c}
        assert r.hist_dict(7) == {} and r._hist is None
        np.testing.assert_array_equal(r.hist, dense)
        assert r.hist is r.hist and r.hist.dtype == np.uint32
    empty = WalkResult(start=z.astype(np.int64), count=z, maxd=z, flags=z, nodes=np.zeros(0, np.int32), hist_packed=np.zeros((50, 0), np.uint16),
                       hist_cols=np.zeros(0, np.int64))
    assert not empty.hist.any() and empty.hist.shape == (50, N_ENTITY_TYPES)
    plain = WalkResult(start=z.astype(np.int64), count=z, maxd=z, flags=z, nodes=np.zeros(0, np.int32))
    assert plain.hist is None
    plain.hist = dense
    assert plain.hist is dense

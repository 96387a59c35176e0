"""CPU-only: the C-ABI library loads and exports every symbol include/abb200.h declares; host-side CSR build is exact."""

from __future__ import annotations

import re
from pathlib import Path

import numpy as np
import pytest

from golden_util import SMALL_FIXTURES, edge_arrays, load, oracle_graph, seeded_graph
from oracle import oracle as orc

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    from agent_bom_b200 import _lib

    header = (ROOT / "include" / "abb200.h").read_text()
    declared = set(re.findall(r"\b(abb_[a-z0-9_]+)\s*\(", header))
    lib = _lib.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"libabb200.so lacks {missing}"
    assert declared == set(_lib.EXPORTS), f"binding table out of sync: {declared ^ set(_lib.EXPORTS)}"
    assert lib.abb_version() == 110


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the engine must refuse loudly (never compute on the CPU)."""
    from agent_bom_b200 import _lib
    from agent_bom_b200.engine import DeviceGraph
    from agent_bom_b200.graph import csr as csrmod

    if _lib.load().abb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    src, dst, rel, flags, nt = seeded_graph(32, 100, 5)
    h = csrmod.from_arrays(None, nt, src, dst, rel, flags)
    with pytest.raises(_lib.EngineUnavailable):
        DeviceGraph.upload(h)


@pytest.mark.parametrize("name", SMALL_FIXTURES + ["estate_dense_40"])
def test_host_csr_build_matches_adjacency_model(name):
    """a1: abb_csr_build_host == numpy restatement of add_edge (container.py:146-198)."""
    from agent_bom_b200.graph import csr as csrmod

    doc = load(name)
    og = oracle_graph(name)
    src, dst, rel, flags = edge_arrays(doc)
    h = csrmod.from_arrays(doc["node_ids"], np.asarray(doc["node_types"], dtype=np.uint8), src, dst, rel, flags, n_real=doc["n_real"])
    for arr in ("fwd_off", "fwd_nbr", "fwd_meta", "fwd_eid", "rev_off", "rev_nbr", "rev_meta", "rev_eid"):
        np.testing.assert_array_equal(getattr(h, arr), getattr(og, arr), err_msg=arr)


def test_host_csr_build_random_and_errors():
    from agent_bom_b200 import _lib
    from agent_bom_b200.graph import csr as csrmod

    for seed in range(4):
        src, dst, rel, flags, nt = seeded_graph(500, 4000, seed)
        h = csrmod.from_arrays(None, nt, src, dst, rel, flags)
        og = orc.build_csr(len(nt), src, dst, rel, flags, nt)
        for arr in ("fwd_off", "fwd_nbr", "fwd_meta", "fwd_eid", "rev_off", "rev_nbr", "rev_meta", "rev_eid"):
            np.testing.assert_array_equal(getattr(h, arr), getattr(og, arr), err_msg=arr)
    with pytest.raises(_lib.AbbError):
        csrmod.build_rows(4, np.asarray([0, 9]), np.asarray([1, 2]), np.asarray([1, 1]), np.asarray([1, 1]))
    empty = csrmod.build_rows(3, np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0))
    assert empty["fwd_off"].tolist() == [0, 0, 0, 0]


def test_walk_spec_mapping():
    """The reference-function -> walk-spec table (include/abb200.h) is what the parity tests assume."""
    from agent_bom_b200 import _lib
    from agent_bom_b200.engine import DeviceGraph as DG

    s = DG.spec_impact_of(4)
    assert (s.direction, s.max_depth, s.rel_mask) == (_lib.DIR_REVERSE, 4, 0xFFFFFFFF) and not (s.flags & _lib.WALK_TRAVERSABLE_ONLY)
    s = DG.spec_bfs(3, True)
    assert s.direction == _lib.DIR_FORWARD and s.flags & _lib.WALK_TRAVERSABLE_ONLY and s.flags & _lib.WALK_PARENTS
    s = DG.spec_traverse(_lib.DIR_BOTH, 2, 10, 20, False, 0, True, False, False)
    assert s.rel_mask == 0xFFFFFFFF & ~((1 << 26) | (1 << 27) | (1 << 28)) and not (s.flags & _lib.WALK_MARK_ROOTS)
    assert (s.max_nodes, s.max_edges) == (10, 20)
    s = DG.spec_traverse(_lib.DIR_BOTH, 2, -1, -1, True, 0b110, False, True, True)
    assert s.rel_mask == 0 and s.flags & _lib.WALK_MARK_ROOTS and s.flags & _lib.WALK_TRAVERSABLE_ONLY
    s = DG.spec_shortest_path()
    assert s.max_depth < 0 and s.flags & _lib.WALK_TARGET


def test_next_row_entry_points_refuse_without_device():
    """The §8(f) entry points (union reduction, lateral search, bottleneck score, backend traversals) have no CPU path either."""
    from agent_bom_b200 import _lib
    from agent_bom_b200.backend import get_backend
    from agent_bom_b200.context_graph import ContextGraph, EdgeKind, GraphEdge, GraphNode, NodeKind
    from agent_bom_b200.effective_reach import annotate_graph
    from agent_bom_b200.engine import group_union
    from agent_bom_b200.lateral import find_lateral_paths

    if _lib.load().abb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    refused = (_lib.AbbError, _lib.EngineUnavailable)
    with pytest.raises(refused):
        group_union([0, 1], [0], [0, 1], [5])
    g = ContextGraph()
    g.add_node(GraphNode("agent:a", NodeKind.AGENT, "a"))
    g.add_node(GraphNode("agent:b", NodeKind.AGENT, "b"))
    g.add_node(GraphNode("vuln:1", NodeKind.VULNERABILITY, "v"))
    g.add_edge(GraphEdge("agent:a", "agent:b", EdgeKind.SHARES_SERVER))
    with pytest.raises(refused):
        find_lateral_paths(g, "agent:a")
    with pytest.raises(refused):
        annotate_graph(g)
    assert "effective_reach" not in g.nodes["vuln:1"].metadata          # nothing was written by a half-run
    b = get_backend("b200")
    b.add_node("a", "agent", "a")
    b.add_edge("a", "b", "uses")
    with pytest.raises(refused):
        b.bfs("a")
    with pytest.raises(refused):
        b.bottleneck_nodes()

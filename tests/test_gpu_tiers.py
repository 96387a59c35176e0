"""Every storage tier of the walk gives the same bits: the parity checks of test_gpu_parity.py re-run with the block tiers
(csrc/walkb.cuh) on, off, one at a time, and shrunk so that queries are handed S1 -> mid -> big -> G1 -> GX on small graphs."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import oracle as orc
from test_gpu_parity import REACH4, assert_slices_equal, graphs_for

pytestmark = pytest.mark.gpu

BIG = 1 << 40        # big_limit: the big block tier takes every forecast-heavy query however many there are
CONFIGS = {
    "default": dict(block_tiers=2, mid_qcap=4096, big_qcap=36864, big_limit=24 * 148),
    "warp-tiers-only": dict(block_tiers=0, mid_qcap=4096, big_qcap=36864, big_limit=BIG),
    "mid-only": dict(block_tiers=1, mid_qcap=4096, big_qcap=36864, big_limit=BIG),
    "big-always": dict(block_tiers=2, mid_qcap=4096, big_qcap=36864, big_limit=BIG),
    "big-never-routed": dict(block_tiers=3, mid_qcap=4096, big_qcap=36864, big_limit=0),
    "all-tiers": dict(block_tiers=3, mid_qcap=4096, big_qcap=36864, big_limit=BIG),
    "shrunk": dict(block_tiers=3, mid_qcap=300, big_qcap=900, big_limit=BIG),       # most queries of the larger graphs overflow twice
}
GRAPHS = [(2000, 12000, 2), (6000, 90000, 3), (30000, 400000, 4), "estate_dense_40"]


@pytest.fixture(params=list(CONFIGS))
def tier_config(request):
    return request.param


def configured(key, name):
    og, dg, nt, rank, h = graphs_for(key)
    for k, v in CONFIGS[name].items():
        dg.set_option(k, v)
        assert dg.get_option(k) == v
    return og, dg


def restore(dg):
    for k, v in CONFIGS["default"].items():
        dg.set_option(k, v)
    dg.set_dedup(True)


@pytest.mark.parametrize("dedup", [True, False])
@pytest.mark.parametrize("key", GRAPHS)
def test_impact_every_tier(key, dedup, tier_config):
    og, dg = configured(key, tier_config)
    try:
        dg.set_dedup(dedup)
        rng = np.random.default_rng(21)
        n = og.n_nodes
        sources = np.concatenate([rng.integers(0, n, size=min(1000, 4 * n)), np.arange(min(n, 64)), [-1, n + 3]]).astype(np.int32)
        for depth in (4, 7):
            want = orc.impact_many(og, sources, depth)
            got = dg.impact_many(sources, depth)
            assert_slices_equal(got, want)
            np.testing.assert_array_equal(got.maxd, want.maxd)
            np.testing.assert_array_equal(got.hist, want.hist)
            np.testing.assert_array_equal(got.flags & 2, want.flags & 2)
    finally:
        restore(dg)


@pytest.mark.parametrize("key", GRAPHS)
def test_bfs_parents_every_tier(key, tier_config):
    og, dg = configured(key, tier_config)
    try:
        rng = np.random.default_rng(22)
        sources = rng.integers(0, og.n_nodes, size=300).astype(np.int32)
        for depth, trav in ((4, True), (6, False)):
            want = orc.bfs_many(og, sources, depth, trav)
            got = dg.bfs_many(sources, depth, trav)
            assert_slices_equal(got, want, gpu_aux="parent", aux_shift=-1)
    finally:
        restore(dg)


@pytest.mark.parametrize("dedup", [True, False])
@pytest.mark.parametrize("key", GRAPHS)
def test_masked_distances_every_tier(key, dedup, tier_config):
    """Unbounded depth, relationship mask, type-filtered emission, depths (the dependency-reach walk)."""
    og, dg = configured(key, tier_config)
    try:
        dg.set_dedup(dedup)
        rng = np.random.default_rng(23)
        sources = rng.integers(0, og.n_nodes, size=250).astype(np.int32)
        for mask, emit in ((REACH4, 0), (0xFFFFFFFF, 0), (REACH4 | (1 << 5) | (1 << 9), 1 << 3)):
            want = orc.distances_many(og, sources, mask)
            spec = dg.spec_distances(mask, emit)
            got = dg.walk(spec, sources)
            if emit:   # the oracle lists every reached node; the device emits only the requested entity types, in the same order
                keep = ((emit >> np.minimum(og.node_type[want.nodes].astype(np.int64), 31)) & 1).astype(bool)
                for q in range(len(sources)):
                    a, b = int(want.off[q]), int(want.off[q + 1])
                    sel = keep[a:b]
                    np.testing.assert_array_equal(got.slice(q), want.nodes[a:b][sel])
                    np.testing.assert_array_equal(got.aux(q, "depth"), want.aux[a:b][sel])
            else:
                assert_slices_equal(got, want, gpu_aux="depth")
    finally:
        restore(dg)


def test_reachable_every_tier(tier_config):
    og, dg = configured((30000, 400000, 4), tier_config)
    try:
        rng = np.random.default_rng(24)
        sources = rng.integers(0, og.n_nodes, size=300).astype(np.int32)
        for depth, trav in ((6, False), (3, True)):
            assert_slices_equal(dg.reachable_many(sources, depth, trav), orc.reachable_many(og, sources, depth, trav))
    finally:
        restore(dg)

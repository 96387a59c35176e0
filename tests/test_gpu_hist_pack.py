"""Entity-type histograms of the host result: the packed form that crosses PCIe (non-zero columns, 16-bit unless a count needs
more; csrc/histpack.cuh) rebuilds exactly the dense table — in numpy (`WalkResult.hist`), in the library
(`abb_walk_result_hist`), with the option off, and when a count does not fit 16 bits."""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from test_gpu_parity import device_graph, graphs_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key", [(6000, 90000, 3), (30000, 400000, 4), "estate_dense_40"])
@pytest.mark.parametrize("zero_copy", [False, True])
def test_packed_histograms_rebuild_the_dense_table(key, zero_copy):
    from agent_bom_b200.engine import DeviceGraph
    from agent_bom_b200.graph.schema import ENTITY_VALUES

    og, dg, nt, rank, h = graphs_for(key)
    rng = np.random.default_rng(77)
    sources = rng.integers(0, og.n_nodes, size=3000).astype(np.int32)
    want = orc.impact_many(og, sources, 4)
    try:
        assert dg.get_option("hist_pack") == 1
        got = dg.walk(DeviceGraph.spec_impact_of(4), sources, zero_copy=zero_copy)
        assert got.hist_packed is not None and got._hist is None
        present = np.flatnonzero(want.hist.any(axis=0))
        np.testing.assert_array_equal(got.hist_cols, present)
        assert got.hist_packed.dtype == (np.uint16 if want.hist.max() < 65536 else np.uint32)
        assert got.hist_packed.shape == (len(sources), len(present))
        for q in (0, 1, 17, 2999):      # the dict form reads the packed rows directly
            assert got.hist_dict(q) == {ENTITY_VALUES[t]: int(c) for t, c in enumerate(want.hist[q]) if c}
        assert got._hist is None
        np.testing.assert_array_equal(got.hist, want.hist)
        dg.set_option("hist_pack", 0)
        dense = dg.walk(DeviceGraph.spec_impact_of(4), sources, zero_copy=zero_copy)
        assert dense.hist_packed is None
        np.testing.assert_array_equal(dense.hist, want.hist)
        assert dense.d2h_bytes > got.d2h_bytes
    finally:
        dg.set_option("hist_pack", 1)
    small = dg.impact_many(sources[:100], 4)          # below the batch threshold the dense table is shipped as before
    assert small.hist_packed is None
    np.testing.assert_array_equal(small.hist, want.hist[:100])


def test_library_side_expansion_and_wide_counts():
    """A hub whose blast radius holds 70 000 nodes of one type: counts need 32 bits; abb_walk_result_hist (the dense ABI
    accessor) rebuilds the same table the numpy path does."""
    from agent_bom_b200 import _lib
    from agent_bom_b200.engine import DeviceGraph, _vp
    from agent_bom_b200.graph.schema import N_ENTITY_TYPES

    leaves = 70_000
    n = leaves + 3
    # leaf -> hub edges: impact_of(hub) walks the reverse direction and reaches every leaf
    src = np.concatenate([np.arange(3, n), [1, 2]]).astype(np.int32)
    dst = np.concatenate([np.zeros(leaves), [0, 1]]).astype(np.int32)
    rel = np.full(len(src), 3, np.int8)
    flags = np.ones(len(src), np.uint8)
    nt = np.full(n, 4, np.uint8); nt[0] = 1; nt[1] = 2; nt[2] = 6
    og = orc.build_csr(n, src, dst, rel, flags, nt)
    dg, _ = device_graph(src, dst, rel, flags, nt, np.arange(n, dtype=np.int32))
    try:
        sources = np.concatenate([np.zeros(8, np.int32), np.arange(1, 1400, dtype=np.int32)])
        want = orc.impact_many(og, sources, 4)
        assert want.hist.max() >= 65536
        got = dg.impact_many(sources, 4)
        assert got.hist_packed is not None and got.hist_packed.dtype == np.uint32
        np.testing.assert_array_equal(got.hist, want.hist)
        # the same through the dense accessor of the C ABI
        lib = _lib.load()
        spec = DeviceGraph.spec_impact_of(4)
        res = _vp()
        with dg._use() as hnd:
            _lib.check(lib.abb_walk_host(hnd, C.byref(spec), sources.ctypes.data, None, None, len(sources), C.byref(res)))
        try:
            mask, width = C.c_uint32(0), C.c_int32(0)
            assert lib.abb_walk_result_hist_packed(res, C.byref(mask), C.byref(width))
            assert width.value == 4 and mask.value == int(sum(1 << int(t) for t in np.flatnonzero(want.hist.any(axis=0))))
            ptr = lib.abb_walk_result_hist(res)
            dense = np.frombuffer((C.c_char * (len(sources) * N_ENTITY_TYPES * 4)).from_address(ptr), dtype=np.uint32).reshape(len(sources), N_ENTITY_TYPES)
            np.testing.assert_array_equal(dense, want.hist)
            assert lib.abb_walk_result_hist(res) == ptr        # built once
        finally:
            lib.abb_walk_result_free(res)
    finally:
        dg.close()

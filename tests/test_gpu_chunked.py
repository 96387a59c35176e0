"""Chunked host walks (csrc/abb200.cu: walk_host_chunked): a large plain batch is walked in pieces (sources assigned by frontier
signature, or contiguous ranges when the walk does not de-duplicate) whose node arenas are copied out while the next piece is
walked; per-query results are put back in caller order.  Same bits as the one-piece call: slices, histograms (packed and dense),
depths, flags, exposure-path rows — with and without de-duplication, whatever the number of pieces."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import oracle as orc
from test_gpu_parity import assert_slices_equal, graphs_for

pytestmark = pytest.mark.gpu


def restore(dg):
    dg.set_option("chunks", 8)
    dg.set_option("chunk_min", 2 << 20)
    dg.set_dedup(True)


def check(got, want):
    assert_slices_equal(got, want)
    np.testing.assert_array_equal(got.maxd, want.maxd)
    np.testing.assert_array_equal(got.hist, want.hist)
    np.testing.assert_array_equal(got.flags & 2, want.flags & 2)


CASES = [((6000, 90000, 3), 2, True), ((6000, 90000, 3), 2, False), ((6000, 90000, 3), 7, True), ((6000, 90000, 3), 3, False),
         ((30000, 400000, 4), 3, True), ((30000, 400000, 4), 2, False), ("estate_dense_40", 7, True), ("estate_dense_40", 2, False)]


@pytest.mark.parametrize("key,chunks,dedup", CASES)
def test_chunked_impact_walk_equals_the_oracle(key, chunks, dedup):
    og, dg, nt, rank, h = graphs_for(key)
    n = og.n_nodes
    rng = np.random.default_rng(100 + chunks)
    size = 2400 + 10 * chunks + int(dedup)             # a batch size no other test uses: the first call has no size hint
    sources = np.concatenate([rng.integers(0, n, size=size - 70), np.arange(min(n, 64)), [-1, n + 3, 0, 0, 5, 5]]).astype(np.int32)
    want = orc.impact_many(og, sources, 4)
    try:
        dg.set_dedup(dedup)
        dg.set_option("chunk_min", 64)
        dg.set_option("chunks", chunks)
        assert dg.get_option("chunks") == chunks and dg.get_option("chunk_min") == 64
        first = dg.impact_many(sources, 4)
        assert dg.get_option("last_host_chunks") == 1            # no hint yet: one piece
        check(first, want)
        second = dg.impact_many(sources, 4)
        assert dg.get_option("last_host_chunks") == chunks
        check(second, want)
        third = dg.walk(dg.spec_impact_of(4), sources, zero_copy=True)     # per-range size hints now exist
        assert dg.get_option("last_host_chunks") == chunks
        check(third, want)
        assert third.hist_packed is not None
        dg.set_option("hist_pack", 0)
        dense = dg.impact_many(sources, 4)
        assert dg.get_option("last_host_chunks") == chunks and dense.hist_packed is None
        check(dense, want)
    finally:
        dg.set_option("hist_pack", 1)
        restore(dg)


@pytest.mark.parametrize("chunks", [2, 5])
def test_chunked_exposure_call(chunks):
    """abb_exposure_host: chunked walk next to the path pipeline on its own stream."""
    og, dg, nt, rank, h = graphs_for("estate_dense_40")
    findings = np.flatnonzero((nt == 8) | (nt == 9)).astype(np.int32)
    findings = np.concatenate([findings, findings[: 11 + chunks]])          # batch size of this test only
    want = orc.impact_many(og, findings, 4)
    rows = dg.exposure_paths_many(findings)
    try:
        dg.set_option("chunk_min", 32)
        dg.set_option("chunks", chunks)
        for call in range(3):
            w, p = dg.exposure_many(findings, 4, zero_copy=bool(call & 1))
            assert dg.get_option("last_host_chunks") == (1 if call == 0 else chunks)
            check(w, want)
            np.testing.assert_array_equal(p.off, rows.off)
            if not call & 1:            # the zero-copy form carries the factorised rows only
                np.testing.assert_array_equal(p.hops, rows.hops)
                np.testing.assert_array_equal(p.rels, rows.rels)
    finally:
        restore(dg)


def test_other_walks_are_not_chunked():
    """Parents / depths / edges / multi-root batches keep the one-piece path whatever the option says."""
    og, dg, nt, rank, h = graphs_for((6000, 90000, 3))
    rng = np.random.default_rng(9)
    sources = rng.integers(0, og.n_nodes, size=777).astype(np.int32)
    try:
        dg.set_option("chunk_min", 64)
        dg.set_option("chunks", 4)
        for _ in range(2):
            want = orc.bfs_many(og, sources, 4, True)
            got = dg.bfs_many(sources, 4, True)
            assert dg.get_option("last_host_chunks") == 1
            assert_slices_equal(got, want, gpu_aux="parent", aux_shift=-1)
    finally:
        restore(dg)

"""Full-size checks through size-independent properties (the oracle is too slow to replay millions of traversals).

Runs on the M2 estate (1 M nodes / 10 M edges, every finding a source) by default; set ABB_SCALE_WORKLOAD=L for the
10 M-node / 107 M-edge estate of BASELINE.json.  Properties:
  * two independent execution paths agree on every source: root-frontier de-duplicated walk == plain walk;
  * determinism: a second run returns identical arrays;
  * monotonicity in depth: reach(d) is a prefix-compatible subset of reach(d+1), counts never shrink;
  * accounting: histogram rows sum to the number of non-ghost reached nodes; max depth <= limit; no source lists itself;
  * exposure-path rows: offsets are monotone, every row ends in its finding, flat rows == links x templates,
    agent/server columns have the right entity types;
  * a random sample is replayed on the CPU oracle bit-for-bit.
"""

from __future__ import annotations

import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

AGENTS = {"M1": 2_000, "M2": 20_700, "L": 215_000}


@pytest.fixture(scope="module")
def scale():
    from agent_bom_b200 import estate
    from agent_bom_b200.engine import DeviceGraph
    from agent_bom_b200.graph import csr as csrmod

    wl = os.environ.get("ABB_SCALE_WORKLOAD", "M2")
    est = estate.generate(AGENTS[wl], 2145, estate.BENCH_KNOBS, exact_rank=False)
    host = csrmod.from_arrays(None, est.node_type, est.src, est.dst, est.rel, est.flags, node_rank=est.node_rank)
    dg = DeviceGraph.upload(host)
    og = orc.OracleGraph(n_nodes=host.n_nodes, fwd_off=host.fwd_off, fwd_nbr=host.fwd_nbr, fwd_meta=host.fwd_meta, fwd_eid=host.fwd_eid,
                         rev_off=host.rev_off, rev_nbr=host.rev_nbr, rev_meta=host.rev_meta, rev_eid=host.rev_eid, node_type=host.node_type)
    yield est, host, dg, og
    dg.close()


def test_dedup_and_plain_paths_agree_on_every_source(scale):
    est, host, dg, og = scale
    # the plain path stores every source's list separately: bound it to a few hundred million result nodes
    stride = 64 if len(est.findings) > 2_000_000 else 1
    f = np.ascontiguousarray(est.findings[::stride])
    dg.set_dedup(True)
    a = dg.impact_many(f, 4)
    stats = dg.last_walk_stats()
    assert 0 < stats["groups"] < len(f)                      # sharing really happened
    a2 = dg.impact_many(f, 4)
    dg.set_dedup(False)
    b = dg.impact_many(f, 4)
    dg.set_dedup(True)
    for x, y in ((a, a2), (a, b)):
        np.testing.assert_array_equal(x.count, y.count)
        np.testing.assert_array_equal(x.maxd, y.maxd)
        np.testing.assert_array_equal(x.hist, y.hist)
        np.testing.assert_array_equal(x.flags, y.flags)
    # slice contents: compare through a per-source checksum (order-sensitive) computed with numpy on both arenas
    def checksums(r):
        # order-sensitive hash per slice, sum((position_in_slice + 1) * node) modulo 2^64, from two cumulative sums over the arena
        starts, counts = r.start, r.count.astype(np.int64)
        pos = np.arange(r.nodes.shape[0], dtype=np.int64)
        c1 = np.concatenate([[0], np.cumsum(r.nodes.astype(np.int64))])
        c2 = np.concatenate([[0], np.cumsum(r.nodes.astype(np.int64) * pos)])
        s, e = starts, starts + counts
        return (c2[e] - c2[s]) - (s - 1) * (c1[e] - c1[s])
    np.testing.assert_array_equal(checksums(a), checksums(b))
    # and a direct slice comparison on a sample
    rng = np.random.default_rng(5)
    for q in rng.integers(0, len(f), size=3000):
        np.testing.assert_array_equal(a.slice(q), b.slice(q))


def test_depth_monotonicity_and_accounting(scale):
    est, host, dg, og = scale
    f = est.findings[:: max(1, len(est.findings) // 200_000)]
    prev = None
    for depth in (1, 2, 3, 4):
        r = dg.impact_many(f, depth)
        assert int(r.maxd.max()) <= depth
        ghost_free = host.node_type[r.nodes] != 255
        # histogram accounting on every source
        c1 = np.concatenate([[0], np.cumsum(ghost_free.astype(np.int64))])
        s, e = r.start, r.start + r.count
        np.testing.assert_array_equal(r.hist.sum(axis=1).astype(np.int64), c1[e] - c1[s])
        if prev is not None:
            assert (r.count >= prev.count).all()
            # BFS order: the shallower result is a prefix of the deeper one
            for q in np.random.default_rng(depth).integers(0, len(f), size=500):
                n = int(prev.count[q])
                np.testing.assert_array_equal(r.slice(q)[:n], prev.slice(q))
        prev = r
    # no source lists itself
    for q in np.random.default_rng(9).integers(0, len(f), size=2000):
        assert f[q] not in set(prev.slice(q).tolist())


def test_exposure_rows_structure(scale):
    est, host, dg, og = scale
    f = np.ascontiguousarray(est.findings[:: (8 if len(est.findings) > 2_000_000 else 1)])   # keep the expanded rows to a few GB on the host
    rows = dg.exposure_paths_many(f)
    assert (np.diff(rows.off) >= 0).all() and int(rows.off[-1]) == rows.hops.shape[0]
    per_finding = np.diff(rows.off)
    np.testing.assert_array_equal(np.repeat(f, per_finding), rows.hops[:, 3])
    t = host.node_type
    assert set(np.unique(t[rows.hops[:, 1]]).tolist()) <= {1}                          # server column
    assert set(np.unique(t[rows.hops[:, 0]]).tolist()) <= {0, 1, 13, 17}               # agent / user / service account, or the server itself
    three_hop = rows.hops[:, 2] < 0
    assert (rows.rels[three_hop, 2] == -2).all() and (rows.rels[~three_hop] >= -1).all()
    # flat rows == links x templates
    link_rows = np.diff(rows.link_row_off)
    assert int(link_rows.sum()) == rows.hops.shape[0]
    tmpl_idx = np.repeat(rows.link_template - rows.link_row_off[:-1], link_rows) + np.arange(rows.hops.shape[0])
    np.testing.assert_array_equal(rows.template[tmpl_idx, 0], rows.hops[:, 0])
    np.testing.assert_array_equal(rows.template[tmpl_idx, 1], rows.hops[:, 1])
    np.testing.assert_array_equal(rows.template[tmpl_idx, 2], rows.ncred)
    np.testing.assert_array_equal(rows.template[tmpl_idx, 3], rows.ntool)


def test_oracle_replay_of_a_sample(scale):
    est, host, dg, og = scale
    rng = np.random.default_rng(17)
    sel = np.sort(rng.choice(est.findings, size=4000, replace=False)).astype(np.int32)
    got_w, got_p = dg.exposure_many(sel, 4)
    want_w = orc.impact_many(og, sel, 4)
    want_p = orc.derived_paths(og, sel, est.node_rank)
    np.testing.assert_array_equal(got_w.count, np.diff(want_w.off).astype(np.int32))
    for q in range(len(sel)):
        a, b = int(want_w.off[q]), int(want_w.off[q + 1])
        np.testing.assert_array_equal(got_w.slice(q), want_w.nodes[a:b])
    np.testing.assert_array_equal(got_w.hist, want_w.hist)
    np.testing.assert_array_equal(got_p.hops, want_p.hops)
    np.testing.assert_array_equal(got_p.rels, want_p.rels)
    np.testing.assert_array_equal(got_p.ncred, want_p.ncred)
    np.testing.assert_array_equal(got_p.ntool, want_p.ntool)


def test_ranked_pages_match_a_host_sort(scale):
    """Device ranking of every exposure-path row vs a numpy stable sort of the expanded rows (api/routes/graph.py:782-786 order)."""
    est, host, dg, og = scale
    stride = 8 if len(est.findings) > 2_000_000 else 1
    f = np.ascontiguousarray(est.findings[::stride])
    rows = dg.exposure_paths_many(f)
    n_rows = rows.hops.shape[0]
    # base risk of a finding = severity rank * 20 (risk_score is 0 in the estates): critical 100, high 80, medium 60, low 40
    base_of_sev = np.asarray([100.0, 80.0, 60.0, 40.0])
    base = base_of_sev[est.node_sev[f]]
    base_vals, base_id = np.unique(base, return_inverse=True)
    scores = np.empty((len(base_vals), 5, 15))
    for bi, b in enumerate(base_vals.tolist()):
        for nc in range(5):
            for nt in range(15):
                risk = b
                risk += min(10.0, nc * 3.0)
                risk += min(10.0, nt * 0.75)
                scores[bi, nc, nt] = round(min(100.0, risk), 2)
    levels = np.unique(scores)
    table = np.searchsorted(levels, scores).astype(np.uint32)
    # labels are unique per node in the estates, so distinct-label counts == edge counts of the server
    ncu = np.zeros(host.n_nodes, dtype=np.int32); ntu = np.zeros(host.n_nodes, dtype=np.int32)
    ncu[rows.hops[:, 1]] = rows.ncred; ntu[rows.hops[:, 1]] = rows.ntool
    per_finding = np.diff(rows.off)
    row_base = np.repeat(base_id, per_finding)
    risk = scores[row_base, np.minimum(rows.ncred, 4), np.minimum(rows.ntool, 14)]
    nh = np.where(rows.hops[:, 2] >= 0, 4, 3)
    order = np.lexsort((np.arange(n_rows), -ntu[rows.hops[:, 1]], -ncu[rows.hops[:, 1]], -nh, -risk))
    for offset, limit in ((0, 20000), (n_rows // 2, 5000), (max(0, n_rows - 3000), 5000)):
        page, rank, total = dg.rank_exposure_paths(f, base_id.astype(np.int32), table, ncu, ntu, offset, limit)
        assert total == n_rows
        want = order[offset: offset + limit]
        np.testing.assert_array_equal(page.hops, rows.hops[want])
        np.testing.assert_array_equal(page.rels, rows.rels[want])
        np.testing.assert_array_equal(levels[rank], risk[want])


def test_L_oracle_replay_including_the_heaviest_sources():
    """L-scale parity where it is hardest: >= 50 K findings of the 10 M-node / 107 M-edge estate — evenly spaced ones plus the 1 000
    with the largest forecast reach (the 30 K-node blast radii that cross every tier hand-off) — replayed on the CPU oracle bit for bit:
    slice order, histograms, max depth, and the exposure-path rows of the same findings."""
    from agent_bom_b200 import estate
    from agent_bom_b200.engine import DeviceGraph
    from agent_bom_b200.graph import csr as csrmod
    from bench import forecast_weight, host_threads

    est = estate.generate(AGENTS["L"], 2145, estate.BENCH_KNOBS, exact_rank=False)
    host = csrmod.from_arrays(None, est.node_type, est.src, est.dst, est.rel, est.flags, node_rank=est.node_rank)
    dg = DeviceGraph.upload(host)
    og = orc.OracleGraph(n_nodes=host.n_nodes, fwd_off=host.fwd_off, fwd_nbr=host.fwd_nbr, fwd_meta=host.fwd_meta, fwd_eid=host.fwd_eid,
                         rev_off=host.rev_off, rev_nbr=host.rev_nbr, rev_meta=host.rev_meta, rev_eid=host.rev_eid, node_type=host.node_type)
    try:
        f = est.findings
        w = forecast_weight(host)[f]
        heavy = f[np.argsort(-w, kind="stable")[:1000]]
        sel = np.unique(np.concatenate([f[:: max(1, len(f) // 50_000)], heavy])).astype(np.int32)
        assert len(sel) >= 50_000
        threads = host_threads()
        # the device walks ALL findings in one batch (the production shape: frontier groups are shared across the whole batch) ...
        got_all = dg.impact_many(f, 4)
        pos = np.searchsorted(f, sel)
        assert np.array_equal(f[pos], sel)
        want = orc.impact_many(og, sel, 4, threads=threads)
        np.testing.assert_array_equal(got_all.count[pos], np.diff(want.off).astype(np.int32))
        np.testing.assert_array_equal(got_all.hist[pos], want.hist)
        np.testing.assert_array_equal(got_all.maxd[pos], want.maxd)
        assert int(np.diff(want.off).max()) > 25_000          # the heavy ones really are in the sample
        for k, q in enumerate(pos.tolist()):
            a, b = int(want.off[k]), int(want.off[k + 1])
            if not np.array_equal(got_all.slice(q), want.nodes[a:b]):
                raise AssertionError(f"finding {sel[k]}: slice differs from the oracle")
        # ... and the selected ones alone, as a caller with a small batch would (different grouping, same bits), with their path rows
        got_w, got_p = dg.exposure_many(sel, 4)
        np.testing.assert_array_equal(got_w.count, np.diff(want.off).astype(np.int32))
        for k in range(0, len(sel), 7):
            a, b = int(want.off[k]), int(want.off[k + 1])
            np.testing.assert_array_equal(got_w.slice(k), want.nodes[a:b])
        want_p = orc.derived_paths(og, sel, est.node_rank, threads=threads)
        np.testing.assert_array_equal(got_p.hops, want_p.hops)
        np.testing.assert_array_equal(got_p.rels, want_p.rels)
        np.testing.assert_array_equal(got_p.ncred, want_p.ncred)
        np.testing.assert_array_equal(got_p.ntool, want_p.ntool)
    finally:
        dg.close()

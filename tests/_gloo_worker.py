"""Worker of tests/test_host_api.py::test_gloo_world2_broadcast_and_shard (launched by torch.distributed.run, backend gloo)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from agent_bom_b200 import dist as abdist  # noqa: E402
from agent_bom_b200 import estate  # noqa: E402
from agent_bom_b200.graph import csr as csrmod  # noqa: E402

info = abdist.init_from_env("gloo")
est = estate.generate(60, 7, estate.Knobs.dense(4, 8, 2))
ref = csrmod.from_arrays(None, est.node_type, est.src, est.dst, est.rel, est.flags, node_rank=est.node_rank)
host = ref if info.rank == 0 else None
findings = est.findings if info.rank == 0 else None
tensors, n, m = abdist.broadcast_csr(host, info, "cpu")
f = abdist.broadcast_array(findings, info, "cpu")
assert (n, m) == (ref.n_nodes, ref.n_entries)
for name in abdist.CSR_FIELDS:
    want = getattr(ref, name)
    want = want.view(np.int32) if want.dtype == np.uint32 else want
    assert np.array_equal(tensors[name].numpy(), want), name
assert np.array_equal(f.numpy(), est.findings)
lo, hi = abdist.shard_bounds(len(f), info.world, info.rank)
assert int(abdist.sum_over_ranks(hi - lo, info, "cpu")) == len(f)
assert abdist.max_over_ranks(float(info.rank), info, "cpu") == float(info.world - 1)
abdist.barrier(info)
sys.stdout.write("rank%d-ok\n" % info.rank)
sys.stdout.flush()

# ── dependency reach with the agents split across ranks: the oracle stands in for each rank's device call ──────────────
from agent_bom_b200.graph.schema import REACH_MASK, VULN_PKG_MASK  # noqa: E402
from oracle import oracle as orc  # noqa: E402

og = orc.build_csr(ref.n_nodes, est.src, est.dst, est.rel, est.flags, est.node_type)
agents = np.flatnonzero(est.node_type == 0).astype(np.int32)


def local(shard):
    return orc.dependency_reach(og, shard, REACH_MASK, VULN_PKG_MASK, est.node_rank)


merged = abdist.dependency_reach_sharded(local, agents, est.node_rank, info, "cpu")
whole = local(agents)
for key in abdist.REACH_KEYS:
    assert np.array_equal(np.asarray(merged[key]), np.asarray(whole[key])), key
assert len(agents) > 10 and int(np.asarray(whole["pkg_off"])[-1]) > 0
abdist.barrier(info)
sys.stdout.write("rank%d-reach-ok\n" % info.rank)
sys.stdout.flush()

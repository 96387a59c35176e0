"""Graph-walk dependency reachability on the device.

Drop-in for the reference's ``compute_dependency_reach``
(``/root/reference/src/agent_bom/graph/dependency_reach.py:109-220``): BFS from
every agent along USES / DEPENDS_ON / CONTAINS / PROVIDES_TOOL (unbounded
depth), per package the sorted reaching agents and the minimum hop count, per
vulnerability the packages it is attached to (AFFECTS / VULNERABLE_TO, either
direction) and the union / minimum over them.  Both passes run in the CUDA
engine (``abb_dependency_reach_host``); this module only turns indices back
into the reference's dataclasses.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .schema import REACH_MASK, VULN_PKG_MASK, EntityType, enum_value


@dataclass(frozen=True)
class PackageReachability:
    package_id: str
    reachable_from: tuple[str, ...]
    min_hop_distance: int

    @property
    def reachable(self) -> bool:
        return bool(self.reachable_from)


@dataclass(frozen=True)
class VulnerabilityReachability:
    vulnerability_id: str
    package_ids: tuple[str, ...]
    reachable_from: tuple[str, ...]
    min_hop_distance: int

    @property
    def reachable(self) -> bool:
        return bool(self.reachable_from)


@dataclass(frozen=True)
class ReachabilityReport:
    packages: dict[str, PackageReachability]
    vulnerabilities: dict[str, VulnerabilityReachability]

    @property
    def reachable_vulnerability_ids(self) -> tuple[str, ...]:
        return tuple(sorted(v.vulnerability_id for v in self.vulnerabilities.values() if v.reachable))


def compute_dependency_reach(graph, *, rank_info=None, collective_device=None) -> ReachabilityReport:
    """``graph`` is an ``agent_bom_b200.graph.UnifiedGraph`` (use ``UnifiedGraph.from_graph`` to adopt a reference graph).

    With ``rank_info`` (``agent_bom_b200.dist.init_from_env()``) of a multi-rank job, every rank holding the same graph, the
    per-agent BFSs are split across ranks and the per-package union / minimum is exchanged with one all-gather
    (``dist.dependency_reach_sharded``); every rank returns the full report."""
    csr = graph.csr
    ids = csr.node_ids
    agents = np.asarray([csr.idx(n.id) for n in graph.nodes.values() if enum_value(n.entity_type) == EntityType.AGENT.value], dtype=np.int32)
    if rank_info is not None and rank_info.world > 1:
        from ..dist import dependency_reach_sharded

        dg = graph.device_graph
        out = dependency_reach_sharded(lambda shard: dg.dependency_reach(shard, REACH_MASK, VULN_PKG_MASK), agents, csr.node_rank, rank_info,
                                       collective_device if collective_device is not None else f"cuda:{graph.device}")
    else:
        out = graph.device_graph.dependency_reach(agents, REACH_MASK, VULN_PKG_MASK)
    packages: dict[str, PackageReachability] = {}
    out = {k: np.asarray(v) for k, v in out.items()}
    po, pa, pm = out["pkg_off"].tolist(), out["pkg_agents"].tolist(), out["pkg_minhop"].tolist()
    for i, p in enumerate(out["pkg_ids"].tolist()):
        packages[ids[p]] = PackageReachability(ids[p], tuple(ids[a] for a in pa[po[i]: po[i + 1]]), pm[i])
    vulnerabilities: dict[str, VulnerabilityReachability] = {}
    vpo, vp, vao, va, vm = out["vuln_poff"].tolist(), out["vuln_pkgs"].tolist(), out["vuln_aoff"].tolist(), out["vuln_agents"].tolist(), out["vuln_minhop"].tolist()
    for i, v in enumerate(out["vuln_ids"].tolist()):
        vulnerabilities[ids[v]] = VulnerabilityReachability(ids[v], tuple(ids[p] for p in vp[vpo[i]: vpo[i + 1]]),
                                                            tuple(ids[a] for a in va[vao[i]: vao[i + 1]]), vm[i])
    return ReachabilityReport(packages=packages, vulnerabilities=vulnerabilities)

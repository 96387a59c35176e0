#!/usr/bin/env python3
"""compute_dependency_reach across ranks on GPUs — the one place the path has a real exchange step (SURVEY §8e, a9 pass 2).

    python profiles/reach_bench.py [--workload L|M2|M1|S]                      # 1 GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/reach_bench.py --workload L

Every rank holds the replicated CSR (one packed NCCL broadcast), walks its contiguous shard of the agent sources on its own
GPU (pass 1: a BFS per agent, no communication), then the ragged per-package / per-vulnerability agent lists are
all-gathered over NCCL and merged (pass 2; reference graph/dependency_reach.py:109-166).  Prints one JSON line: time of
the sharded call (barrier on both sides, max over ranks), time of the unsplit single-GPU call on rank 0, and whether the
merged answer equals the unsplit device answer (bit for bit, every array)."""

from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="L")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch

    from agent_bom_b200 import dist as abdist, estate
    from agent_bom_b200.engine import DeviceGraph
    from agent_bom_b200.graph import csr as csrmod
    from agent_bom_b200.graph.schema import REACH_MASK, VULN_PKG_MASK
    from bench import WORKLOADS

    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/abb200_nccl.%h.%p.log")
    info = abdist.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    torch.cuda.set_device(info.local_rank)
    device = torch.device("cuda", info.local_rank)
    host = est = None
    if info.rank == 0:
        est = estate.generate(WORKLOADS[args.workload][0], 2145, estate.BENCH_KNOBS, exact_rank=False)
        host = csrmod.from_arrays(None, est.node_type, est.src, est.dst, est.rel, est.flags, node_rank=est.node_rank)
    abdist.warm_up(info, device)
    tensors, n_nodes, n_entries, bstats = abdist.broadcast_csr(host, info, device, return_stats=True)
    agents_t = abdist.broadcast_array(est.agent_nodes if info.rank == 0 else None, info, device)
    agents = agents_t.cpu().numpy()
    node_rank = tensors["node_rank"].cpu().numpy()
    dg = DeviceGraph.adopt(tensors, n_nodes, n_entries, info.local_rank)
    local = lambda shard: dg.dependency_reach(shard, REACH_MASK, VULN_PKG_MASK)
    times = []
    phases = []
    merged = None
    for _ in range(args.reps + 1):
        torch.cuda.synchronize(); abdist.barrier(info)
        t0 = time.perf_counter()
        st = {}
        merged = abdist.dependency_reach_sharded(local, agents, node_rank, info, device, stats=st)
        torch.cuda.synchronize(); abdist.barrier(info)
        times.append(time.perf_counter() - t0)
        phases.append(st)
    sharded_s = abdist.max_over_ranks(min(times[1:]), info, device)
    if info.rank == 0:
        t0 = time.perf_counter()
        whole = local(agents)
        unsplit_s = time.perf_counter() - t0
        same = all(np.array_equal(np.asarray(merged[k]), np.asarray(whole[k])) for k in abdist.REACH_KEYS)
        line = {"what": "compute_dependency_reach, agent BFSs sharded across ranks + all-gather merge", "workload": args.workload, "n_gpus": info.world,
                "agents": int(len(agents)), "nodes": n_nodes, "packages": int(len(whole["pkg_ids"])), "vulnerabilities": int(len(whole["vuln_ids"])),
                "reach_pairs": int(len(whole["pkg_agents"])), "sharded_s": sharded_s, "unsplit_single_gpu_s": unsplit_s,
                "agents_per_s": len(agents) / sharded_s, "equals_unsplit_device_answer": bool(same), "csr_broadcast": bstats,
                "rank0_phases_s": {k: round(v, 4) for k, v in phases[int(np.argmin(times[1:])) + 1].items()},
                "vulnerability_pairs": int(len(whole["vuln_agents"]))}
        print(json.dumps(line), flush=True)
        if not same:
            return 1
    abdist.barrier(info)
    if info.world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())

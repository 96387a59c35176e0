"""Multi-GPU plumbing: one process per GPU, one-time NCCL broadcast of the shared CSR, contiguous source shards.

The exposure graph is replicated (≈1.9 GB at 10 M nodes / 100 M edges, trivial
next to 180 GB of HBM3e); finding sources are independent, so after the
broadcast there is NO data-path collective — each rank traverses its own
contiguous slice of the source list.  ``torch.distributed``
(backend ``nccl`` on GPUs, ``gloo`` in the CPU tests) is only the transport.
"""

from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

CSR_FIELDS = ("fwd_off", "fwd_nbr", "fwd_meta", "fwd_eid", "rev_off", "rev_nbr", "rev_meta", "rev_eid", "node_type", "node_rank")
_TORCH_DTYPE = {"fwd_off": torch.int32, "fwd_nbr": torch.int32, "fwd_meta": torch.uint8, "fwd_eid": torch.int32, "rev_off": torch.int32,
                "rev_nbr": torch.int32, "rev_meta": torch.uint8, "rev_eid": torch.int32, "node_type": torch.uint8, "node_rank": torch.int32}


@dataclass
class RankInfo:
    rank: int
    world: int
    local_rank: int


def init_from_env(backend: str | None = None) -> RankInfo:
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun); no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29517")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return RankInfo(rank, world, local)


def shard_bounds(n_items: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first ``n_items % world`` ranks get one extra item."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(items, info: RankInfo):
    lo, hi = shard_bounds(len(items), info.world, info.rank)
    return items[lo:hi]


def _as_tensor(name: str, arr: np.ndarray) -> torch.Tensor:
    a = np.ascontiguousarray(arr)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a)


def warm_up(info: RankInfo, device: torch.device | str) -> None:
    """Create the communicator before anything is timed (NCCL initialises lazily on the first collective)."""
    if info.world > 1:
        t = torch.zeros(1, dtype=torch.int32, device=torch.device(device))
        dist.all_reduce(t)
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize()


def broadcast_csr(host_csr, info: RankInfo, device: torch.device | str, return_stats: bool = False):
    """Rank 0 holds ``host_csr`` (others pass None); returns (device tensors, n_nodes, n_entries[, stats]) on every rank.

    The ten CSR arrays travel as ONE packed byte buffer (256-byte aligned fields): rank 0 uploads into its slices, a
    single broadcast replicates the buffer — over NVLink 5 / NVSwitch with the ``nccl`` backend — and every rank views
    the fields in place.  ``stats`` times the upload and the collective separately (the collective with CUDA events).
    """
    device = torch.device(device)
    meta = torch.zeros(2, dtype=torch.int64)
    if info.rank == 0:
        meta[0], meta[1] = host_csr.n_nodes, host_csr.n_entries
    if info.world > 1:
        m = meta.to(device) if device.type == "cuda" else meta
        dist.broadcast(m, src=0)
        meta = m.cpu()
    n_nodes, n_entries = int(meta[0]), int(meta[1])
    lengths = {"fwd_off": n_nodes + 1, "rev_off": n_nodes + 1, "node_type": n_nodes, "node_rank": n_nodes}
    layout, total = {}, 0
    for name in CSR_FIELDS:
        length = lengths.get(name, n_entries)
        nbytes = length * torch.empty(0, dtype=_TORCH_DTYPE[name]).element_size()
        layout[name] = (total, nbytes, length)
        total += (nbytes + 255) & ~255
    buf = torch.empty(max(total, 256), dtype=torch.uint8, device=device)
    tensors = {name: buf[off: off + nbytes].view(_TORCH_DTYPE[name]) for name, (off, nbytes, _length) in layout.items()}
    import time as _time

    t0 = _time.perf_counter()
    if info.rank == 0:
        for name in CSR_FIELDS:
            src = _as_tensor(name, getattr(host_csr, name))
            assert src.numel() == layout[name][2], (name, src.numel(), layout[name][2])
            tensors[name].copy_(src)
    if device.type == "cuda":
        torch.cuda.synchronize()
    upload_s = _time.perf_counter() - t0
    bcast_ms = 0.0
    if info.world > 1:
        if device.type == "cuda":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            dist.broadcast(buf, src=0)
            e1.record()
            torch.cuda.synchronize()
            bcast_ms = e0.elapsed_time(e1)
        else:
            t1 = _time.perf_counter()
            dist.broadcast(buf, src=0)
            bcast_ms = 1000.0 * (_time.perf_counter() - t1)
    if return_stats:
        stats = {"bytes": int(total), "rank0_upload_s": round(upload_s, 4), "broadcast_ms": round(bcast_ms, 3),
                 "broadcast_GBps": round(total / 1e9 / (bcast_ms / 1000.0), 1) if bcast_ms > 0 else None, "collectives": 1 if info.world > 1 else 0}
        return tensors, n_nodes, n_entries, stats
    return tensors, n_nodes, n_entries


def broadcast_array(arr: np.ndarray | None, info: RankInfo, device: torch.device | str, dtype=torch.int32) -> torch.Tensor:
    """Broadcast a 1-D integer array (e.g. the finding-source list) from rank 0."""
    device = torch.device(device)
    n = torch.zeros(1, dtype=torch.int64)
    if info.rank == 0:
        n[0] = len(arr)
    if info.world > 1:
        nn = n.to(device) if device.type == "cuda" else n
        dist.broadcast(nn, src=0)
        n = nn.cpu()
    if info.rank == 0:
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(dtype).to(device)
    else:
        t = torch.empty(int(n[0]), dtype=dtype, device=device)
    if info.world > 1:
        dist.broadcast(t, src=0)
    return t


def max_over_ranks(value: float, info: RankInfo, device: torch.device | str) -> float:
    if info.world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, info: RankInfo, device: torch.device | str) -> float:
    if info.world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_floats(values, info: RankInfo, device: torch.device | str) -> list[list[float]]:
    """Every rank's list of floats, by rank (diagnostics: per-rank kernel times)."""
    if info.world == 1:
        return [[float(v) for v in values]]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    parts = [torch.zeros_like(t) for _ in range(info.world)]
    dist.all_gather(parts, t)
    return [[float(x) for x in p.cpu().tolist()] for p in parts]


def barrier(info: RankInfo) -> None:
    if info.world > 1:
        dist.barrier()


# ── dependency reach across ranks: the one place the path has a real exchange step (SURVEY.md §8e, a9 pass 2) ───────────
REACH_KEYS = ("pkg_ids", "pkg_off", "pkg_agents", "pkg_minhop", "vuln_ids", "vuln_poff", "vuln_pkgs", "vuln_aoff", "vuln_agents", "vuln_minhop")


def all_gather_ragged(arr: np.ndarray, info: RankInfo, device: torch.device | str, keep_on_device: bool = False) -> list:
    """Every rank's 1-D array on every rank: one size exchange, one padded all-gather (NCCL over NVLink on GPUs, gloo on CPU).
    ``keep_on_device`` returns the parts as tensors on ``device`` (for a merge that runs there) instead of numpy arrays."""
    a = np.ascontiguousarray(arr)
    device = torch.device(device)
    if info.world == 1:
        return [torch.from_numpy(a).to(device)] if keep_on_device else [a]
    size = torch.tensor([a.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(info.world)]
    dist.all_gather(sizes, size)
    counts = [int(s.item()) for s in sizes]
    width = max(max(counts), 1)
    buf = torch.zeros(width, dtype=torch.from_numpy(a[:0]).dtype, device=device)
    if a.shape[0]:
        buf[: a.shape[0]] = torch.from_numpy(a).to(device)
    parts = [torch.empty_like(buf) for _ in range(info.world)]
    dist.all_gather(parts, buf)
    if keep_on_device:
        return [p[:c] for p, c in zip(parts, counts)]
    return [p[:c].cpu().numpy() for p, c in zip(parts, counts)]


def _host(x) -> np.ndarray:
    return x.cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def merge_dependency_reach(partials: list[dict], node_rank, device: torch.device | str | None = None) -> dict:
    """Combine per-rank results of ``DeviceGraph.dependency_reach`` computed over disjoint agent shards.

    Package and vulnerability tables are graph-constant (identical on every rank); what differs is who reaches what:
    per package / vulnerability the agent lists are concatenated and re-sorted by id-string rank (the reference's
    ``tuple(sorted(...))``, graph/dependency_reach.py:139-145,156-164), and the hop minimum is taken over the ranks that
    reach it at all (a rank that does not reach it reports 0, which must not win the minimum).

    On a CUDA ``device`` the (group, rank) sort of the concatenated pairs — hundreds of millions on a 10 M-node estate — runs
    there as one 64-bit key sort; the numpy path (``device`` None / cpu) is the same computation for the gloo tests."""
    node_rank = np.asarray(node_rank)
    first = partials[0]
    out = {k: first[k] for k in ("pkg_ids", "vuln_ids", "vuln_poff", "vuln_pkgs")}
    on_gpu = device is not None and torch.device(device).type == "cuda"
    if on_gpu:
        dev = torch.device(device)
        rank_t = torch.from_numpy(np.ascontiguousarray(node_rank, dtype=np.int64)).to(dev)
        span = int(node_rank.max()) + 1 if node_rank.size else 1
    for key_off, key_items, key_min, n_groups in (("pkg_off", "pkg_agents", "pkg_minhop", len(first["pkg_ids"])),
                                                  ("vuln_aoff", "vuln_agents", "vuln_minhop", len(first["vuln_ids"]))):
        big = np.full(n_groups, np.iinfo(np.int32).max, dtype=np.int64)
        all_counts = []
        for p in partials:
            counts = np.diff(_host(p[key_off]).astype(np.int64))
            all_counts.append(counts)
            reached = counts > 0
            big[reached] = np.minimum(big[reached], _host(p[key_min]).astype(np.int64)[reached])
        total = np.sum(all_counts, axis=0) if all_counts else np.zeros(n_groups, np.int64)
        off = np.zeros(n_groups + 1, dtype=np.int64)
        off[1:] = np.cumsum(total)
        if on_gpu:
            keys, items = [], []
            gid = torch.arange(n_groups, dtype=torch.int64, device=dev)
            for p, counts in zip(partials, all_counts):
                itm = p[key_items].to(dev) if torch.is_tensor(p[key_items]) else torch.from_numpy(np.ascontiguousarray(p[key_items], dtype=np.int32)).to(dev)
                grp = torch.repeat_interleave(gid, torch.from_numpy(counts).to(dev), output_size=int(itm.shape[0]))
                keys.append(grp * span + rank_t[itm.long()])
                items.append(itm)
            if items and sum(int(i.shape[0]) for i in items):
                order = torch.argsort(torch.cat(keys))
                merged = torch.cat(items)[order].cpu().numpy()
            else:
                merged = np.zeros(0, np.int32)
        else:
            grp = np.concatenate([np.repeat(np.arange(n_groups, dtype=np.int64), c) for c in all_counts]) if all_counts else np.zeros(0, np.int64)
            itm = np.concatenate([_host(p[key_items]).astype(np.int32, copy=False) for p in partials]) if partials else np.zeros(0, np.int32)
            merged = itm[np.lexsort((node_rank[itm], grp))]
        out[key_off], out[key_items] = off, merged
        out[key_min] = np.where(big == np.iinfo(np.int32).max, 0, big).astype(np.int32)
    return out


def dependency_reach_sharded(local_reach, agents, node_rank, info: RankInfo, device: torch.device | str, stats: dict | None = None) -> dict:
    """``compute_dependency_reach`` with the agent BFSs split across ranks.

    ``local_reach(agent_shard) -> dict`` is this rank's device call (``lambda a: dg.dependency_reach(a, REACH_MASK,
    VULN_PKG_MASK)``).  Pass 1 (a BFS per agent) needs no communication; pass 2's per-package union / minimum is the
    exchange: every rank's ragged agent lists are all-gathered and merged, so every rank ends with the full answer.
    On GPUs the gathered lists stay on the device and are merged there.  ``stats`` (optional) receives the wall-clock seconds of
    the three phases on this rank: ``local_s``, ``gather_s``, ``merge_s``."""
    import time

    def mark():
        if torch.device(device).type == "cuda":
            torch.cuda.synchronize(device)
        return time.perf_counter()

    agents = np.ascontiguousarray(agents, dtype=np.int32)
    lo, hi = shard_bounds(len(agents), info.world, info.rank)
    t0 = mark()
    mine = local_reach(agents[lo:hi])
    t1 = mark()
    if info.world == 1:
        if stats is not None:
            stats.update(local_s=t1 - t0, gather_s=0.0, merge_s=0.0)
        return {k: np.asarray(mine[k]) for k in REACH_KEYS}
    on_gpu = torch.device(device).type == "cuda"
    gathered = {k: all_gather_ragged(np.asarray(mine[k]), info, device, keep_on_device=on_gpu and k in ("pkg_agents", "vuln_agents"))
                for k in ("pkg_off", "pkg_agents", "pkg_minhop", "vuln_aoff", "vuln_agents", "vuln_minhop")}
    t2 = mark()
    partials = []
    for r in range(info.world):
        p = {k: np.asarray(mine[k]) for k in ("pkg_ids", "vuln_ids", "vuln_poff", "vuln_pkgs")}
        p.update({k: gathered[k][r] for k in gathered})
        partials.append(p)
    out = merge_dependency_reach(partials, node_rank, device)
    t3 = mark()
    if stats is not None:
        stats.update(local_s=t1 - t0, gather_s=t2 - t1, merge_s=t3 - t2)
    return out

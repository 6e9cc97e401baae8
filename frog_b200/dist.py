"""Multi-GPU plumbing: shard image pairs across ranks, gather compacted match lists to rank 0.

Image pairs are independent units (match.cpp:638-652 runs them as an OpenMP `for`), so the
data path needs no collective: every rank matches its own pairs.  The only exchange is the
result hand-off -- per-pair counts plus the compacted (first, second) lists -- which goes
GPU-to-GPU over NVLink (NCCL send/recv) to rank 0, mirroring the single writer of
match.cpp:660-745.  The same code runs on CPU tensors over gloo in the tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_pairs(weights, world: int):
    """Longest-processing-time-first assignment of pairs to ranks.  weights[p] = N_first * N_second.
    Returns a list (per rank) of sorted global pair ids."""
    order = np.argsort(-np.asarray(weights, np.float64), kind="stable")
    load = np.zeros(world)
    shards = [[] for _ in range(world)]
    for p in order:
        r = int(np.argmin(load))
        shards[r].append(int(p))
        load[r] += weights[p]
    return [sorted(s) for s in shards]


class DeviceArray:
    """Zero-copy view of a raw CUDA allocation for torch.as_tensor (CUDA array interface v3)."""

    def __init__(self, ptr: int, n_elems: int, typestr: str = "<u4"):
        self.__cuda_array_interface__ = {"shape": (n_elems,), "typestr": typestr, "data": (ptr, False),
                                         "version": 3, "strides": None}


def as_torch_u32(ptr: int, n_elems: int, device) -> torch.Tensor:
    """int32 tensor aliasing `n_elems` uint32 at device address `ptr` (NCCL moves bytes; int32 is
    the widest-supported 4-byte dtype)."""
    if n_elems == 0 or not ptr:
        return torch.zeros(0, dtype=torch.int32, device=device)
    return torch.as_tensor(DeviceArray(ptr, n_elems, "<i4"), device=device)


def gather_match_lists(counts: torch.Tensor, pairs: torch.Tensor, dst: int = 0):
    """counts: [n_local_pairs] int32, pairs: [2 * total_local] int32, both on this rank's device.
    Returns on `dst`: (list of counts tensors per rank, list of pairs tensors per rank); elsewhere None.
    One small all_gather for the sizes, then point-to-point transfers of exactly the bytes needed."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return [counts], [pairs]
    sizes = torch.tensor([counts.numel(), pairs.numel()], dtype=torch.int64, device=counts.device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = torch.stack(all_sizes).cpu().tolist()
    if rank == dst:
        out_counts, out_pairs, ops = [], [], []
        for r in range(world):
            if r == dst:
                out_counts.append(counts)
                out_pairs.append(pairs)
                continue
            c = torch.empty(all_sizes[r][0], dtype=counts.dtype, device=counts.device)
            p = torch.empty(all_sizes[r][1], dtype=pairs.dtype, device=pairs.device)
            out_counts.append(c)
            out_pairs.append(p)
            if c.numel():
                ops.append(dist.P2POp(dist.irecv, c, r))
            if p.numel():
                ops.append(dist.P2POp(dist.irecv, p, r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out_counts, out_pairs
    ops = []
    if counts.numel():
        ops.append(dist.P2POp(dist.isend, counts, dst))
    if pairs.numel():
        ops.append(dist.P2POp(dist.isend, pairs, dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return None


class FixedGather:
    """Gather of compacted match lists to rank `dst` WITHOUT a size exchange, so that no host
    synchronisation sits between the matcher's kernels and the transfer: every rank sends its
    per-pair counts (n_pairs int32) and its list buffer at full capacity (`cap_elems` int32 =
    2 x outer-loop rows: a row yields at most one match, match.cpp:320-327); the receiver reads the
    first 2 * sum(counts) entries.  `slots` receive buffers per peer let a transfer overlap the
    next step.  start() returns the in-flight works; Work.wait() orders the current stream (NCCL)
    or blocks the host (gloo) behind the transfer."""

    def __init__(self, n_pairs: int, cap_elems: int, device, slots: int = 2, dst: int = 0):
        self.world, self.rank, self.dst = dist.get_world_size(), dist.get_rank(), dst
        self.n_pairs, self.cap = n_pairs, cap_elems
        self.counts, self.pairs = [], []
        if self.rank == dst:
            for _ in range(slots):
                self.counts.append([torch.zeros(n_pairs, dtype=torch.int32, device=device) for _ in range(self.world)])
                self.pairs.append([torch.zeros(cap_elems, dtype=torch.int32, device=device) for _ in range(self.world)])

    def start(self, counts: torch.Tensor, pairs_cap: torch.Tensor, slot: int = 0, per_peer: bool = False):
        """Queue the transfer.  Returns the in-flight works: one flat list, or with per_peer=True a dict
        {peer: works} on dst (one batch per peer, so a peer's lists can be consumed -- e.g. copied to the
        host -- while the next peer's are still arriving) and {dst: works} on the senders."""
        assert counts.numel() == self.n_pairs and pairs_cap.numel() == self.cap
        if self.rank != self.dst:
            works = dist.batch_isend_irecv([dist.P2POp(dist.isend, counts, self.dst), dist.P2POp(dist.isend, pairs_cap, self.dst)])
            return {self.dst: works} if per_peer else works
        peers = [r for r in range(self.world) if r != self.dst]
        if per_peer:
            return {r: dist.batch_isend_irecv([dist.P2POp(dist.irecv, self.counts[slot][r], r),
                                               dist.P2POp(dist.irecv, self.pairs[slot][r], r)]) for r in peers}
        ops = []
        for r in peers:
            ops.append(dist.P2POp(dist.irecv, self.counts[slot][r], r))
            ops.append(dist.P2POp(dist.irecv, self.pairs[slot][r], r))
        return dist.batch_isend_irecv(ops) if ops else []

    def received(self, slot: int, r: int):
        """On dst, after the works completed: (counts, pairs[: 2 * sum(counts)]) of rank r != dst."""
        c = self.counts[slot][r]
        return c, self.pairs[slot][r][: 2 * int(c.sum().item())]


def assemble(shards, counts_per_rank, pairs_per_rank, n_pairs: int):
    """Rank 0: undo the sharding.  Returns a list of [m,2] uint32 arrays in global pair order."""
    out = [None] * n_pairs
    for r, ids in enumerate(shards):
        c = counts_per_rank[r].cpu().numpy().astype(np.int64)
        p = pairs_per_rank[r].cpu().numpy().view(np.uint32).reshape(-1, 2)
        off = np.concatenate([[0], np.cumsum(c)])
        for k, pid in enumerate(ids):
            out[pid] = p[off[k]:off[k + 1]]
    return out

"""CPU, world_size 2 over gloo: the N > 1 path of bench.py / frog_b200.dist -- image pairs are
sharded across ranks, every rank produces its own compacted lists, rank 0 gathers exactly the
bytes needed and undoes the sharding."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from frog_b200 import dist as fdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_lists(n_pairs):
    rng = np.random.default_rng(11)
    return [rng.integers(0, 1000, size=(int(rng.integers(0, 40)), 2)).astype(np.uint32) for _ in range(n_pairs)]


def _worker(rank, world, port, n_pairs, weights, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards = fdist.shard_pairs(weights, world)
    truth = _fake_lists(n_pairs)
    mine = [truth[p] for p in shards[rank]]
    counts = torch.tensor([len(l) for l in mine], dtype=torch.int32)
    flat = np.concatenate(mine).reshape(-1) if mine else np.zeros(0, np.uint32)
    pairs = torch.from_numpy(flat.view(np.int32).copy())
    got = fdist.gather_match_lists(counts, pairs, 0)
    if rank == 0:
        lists = fdist.assemble(shards, got[0], got[1], n_pairs)
        q.put(all(np.array_equal(a, b) for a, b in zip(lists, truth)))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_is_partition_and_balanced():
    w = np.array([5, 1, 1, 1, 4, 4, 2, 2, 9, 3], np.float64)
    shards = fdist.shard_pairs(w, 3)
    assert sorted(sum(shards, [])) == list(range(10))
    loads = [w[s].sum() for s in shards]
    assert max(loads) - min(loads) <= w.max()


def test_gather_world2_gloo():
    n_pairs = 13
    weights = list(np.random.default_rng(1).uniform(1, 5, n_pairs))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, weights, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _fixed_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_pairs, cap = 6, 2 * 300
    fg = fdist.FixedGather(n_pairs, cap, torch.device("cpu"), slots=2)
    ok = True
    for step in range(3):  # slots are reused: step 2 lands in slot 0 again
        rng = np.random.default_rng(100 * step + rank)
        lists = [rng.integers(0, 1000, size=(int(rng.integers(0, 40)), 2)).astype(np.uint32) for _ in range(n_pairs)]
        counts = torch.tensor([len(l) for l in lists], dtype=torch.int32)
        buf = torch.full((cap,), -1, dtype=torch.int32)  # capacity buffer: only the head is meaningful
        flat = np.concatenate(lists).reshape(-1).view(np.int32)
        buf[: flat.size] = torch.from_numpy(flat.copy())
        for w in fg.start(counts, buf, step % 2):
            w.wait()
        if rank == 0:
            rng1 = np.random.default_rng(100 * step + 1)
            want = [rng1.integers(0, 1000, size=(int(rng1.integers(0, 40)), 2)).astype(np.uint32) for _ in range(n_pairs)]
            c, p = fg.received(step % 2, 1)
            ok &= c.tolist() == [len(l) for l in want]
            ok &= np.array_equal(p.numpy().view(np.uint32).reshape(-1, 2), np.concatenate(want))
    if rank == 0:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_fixed_capacity_gather_world2_gloo():
    """The size-exchange-free gather bench.py uses for N > 1 (no host synchronisation on the data path)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fixed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True

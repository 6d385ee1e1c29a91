"""Host logic of the multi-GPU modes on CPU: world sharding, slab partition, and the neighbour
exchange protocol of bullet3_b200/slab.py run over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bullet3_b200 import slab


def test_shard_worlds_covers_everything():
    for nw, ws in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        sh = slab.shard_worlds(nw, ws)
        assert len(sh) == ws and sum(n for _, n in sh) == nw
        assert all(sh[i][0] + sh[i][1] == sh[i + 1][0] for i in range(ws - 1))
        assert max(n for _, n in sh) - min(n for _, n in sh) <= 1


def test_slab_partition_balanced_and_ordered():
    rng = np.random.default_rng(0)
    x = rng.uniform(-50, 50, 10007)
    for ws in (1, 2, 4, 8):
        b = slab.slab_boundaries(x, ws)
        s = slab.slab_of(x, b)
        cnt = np.bincount(s, minlength=ws)
        assert cnt.sum() == len(x) and cnt.max() - cnt.min() <= 1
        assert all(x[s == r].max() <= x[s == r + 1].min() for r in range(ws - 1))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


REC = 16  # bytes per record in the CPU protocol test (id, x as two int32/float32 pairs)


def _worker(rank, ws, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        rng = np.random.default_rng(5)  # the same scene on every rank
        n = 4000
        x = rng.uniform(-40, 40, n)
        half = rng.uniform(0.3, 0.8, n)
        bnd = slab.slab_boundaries(x, ws)
        owner = slab.slab_of(x, bnd)
        margin = 1.5
        lo, hi = bnd[rank], bnd[rank + 1]
        mine = np.nonzero(owner == rank)[0]
        left, right = slab.band_masks(x[mine] - half[mine], x[mine] + half[mine], lo, hi, margin)
        sides = [s for s, has in (("left", rank > 0), ("right", rank < ws - 1)) if has]
        peer = {"left": rank - 1, "right": rank + 1}
        cap = n
        send = {s: torch.zeros(cap * REC, dtype=torch.uint8) for s in sides}
        recv = {s: torch.zeros(cap * REC, dtype=torch.uint8) for s in sides}
        cs = {s: torch.zeros(1, dtype=torch.int32) for s in sides}
        cr = {s: torch.zeros(1, dtype=torch.int32) for s in sides}
        counts = {}
        for s in sides:
            ids = mine[left if s == "left" else right]
            rec = np.zeros((len(ids), 4), np.float32)
            rec[:, 0] = ids
            rec[:, 1] = x[ids]
            send[s][: len(ids) * REC] = torch.from_numpy(rec.view(np.uint8).reshape(-1))
            counts[s] = len(ids)
        rc = slab.exchange_buffers(sides, peer, counts, send, recv, cs, cr, REC)
        ok = True
        for s in sides:
            got = recv[s][: rc[s] * REC].numpy().view(np.float32).reshape(-1, 4)
            # what the neighbour must have sent, recomputed here from the shared scene
            p = peer[s]
            theirs = np.nonzero(owner == p)[0]
            l2, r2 = slab.band_masks(x[theirs] - half[theirs], x[theirs] + half[theirs], bnd[p], bnd[p + 1], margin)
            want = theirs[r2 if s == "left" else l2]
            ok &= np.array_equal(np.sort(got[:, 0].astype(np.int64)), np.sort(want))
            ok &= bool(np.allclose(got[:, 1], x[got[:, 0].astype(np.int64)], atol=1e-5))
            # every body of the neighbour that could touch one of mine within the margin is mirrored
            reach = (x[theirs] + half[theirs] >= lo - 0.0) if s == "left" else (x[theirs] - half[theirs] <= hi + 0.0)
            ok &= set(theirs[reach]).issubset(set(want))
        # batched-worlds mode: every rank steps its shard, the totals add up without a data-path collective
        first, cnt = slab.shard_worlds(37, ws)[rank]
        t = torch.tensor([cnt], dtype=torch.int64)
        dist.all_reduce(t)
        ok &= int(t.item()) == 37
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_neighbour_exchange_over_gloo_world_size_2():
    ws = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(ws, port, ret), nprocs=ws, join=True)
        assert dict(ret) == {0: True, 1: True}


@pytest.mark.timeout(180)
def test_neighbour_exchange_over_gloo_world_size_3():
    """a middle rank has two neighbours"""
    ws = 3
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(ws, port, ret), nprocs=ws, join=True)
        assert dict(ret) == {0: True, 1: True, 2: True}

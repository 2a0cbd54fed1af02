"""world_size-2 gloo test of the only collective on the path: the logits all-gather over a
stream-sharded batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from continual_skeletons_b200 import LogitGather, all_gather_logits, any_rank, shard_range


def test_shard_range_partitions():
    for n in (1, 2, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_streams, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_streams * 5, dtype=torch.float32).view(n_streams, 5)
    lo, hi = shard_range(n_streams, rank, world)
    got = all_gather_logits(full[lo:hi].clone(), n_streams)
    ok = bool(torch.equal(got, full))
    # the pre-allocated serving-loop form of the same collective: two launches (buffers alternate), uneven shards compacted
    g = LogitGather(n_streams, 5, "cpu")
    for k in (1.0, 2.0):
        g.launch(full[lo:hi] * k)
        ok = ok and bool(torch.equal(g.result(), full * k))
    # a loop whose length is decided by something rank-local around a body with a collective (bench.py's prewarm): rank 1
    # wants one more batch than rank 0 -- both must run the same number, or the collectives below would pair up wrongly
    want, ran = 2 + rank, 0
    while any_rank(ran < want):
        t = torch.tensor([float(rank)])
        dist.all_reduce(t)
        ran += 1
    ok = ok and ran == 3 and any_rank(False) is False and any_rank(rank == 1) is True
    ret[rank] = ok
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_all_gather_logits_gloo_world2():
    for n_streams in (6, 7):  # even and uneven shards
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), n_streams, ret), nprocs=2, join=True)
        assert ret[0] and ret[1]


def test_all_gather_is_identity_without_process_group():
    x = torch.rand(3, 4)
    assert all_gather_logits(x, 3) is x


def test_logit_gather_single_process_passthrough():
    g = LogitGather(3, 4, "cpu")
    x = torch.rand(3, 4)
    g.launch(x)
    assert g.result() is x


def test_any_rank_without_process_group():
    assert any_rank(True) is True and any_rank(False) is False

"""world_size-2 checks of the multi-process host logic on CPU (gloo): packet sharding, disjoint Philox stream blocks,
the hand-over of the communicator id, and that a sum over ranks of per-rank partial results equals the single-rank
result (what skg_allreduce_results does on the device with NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common
from skirt_b200.parallel import shard_packets, share_unique_id


def test_shard_packets_covers_the_budget():
    for packages, n in ((1e6, 1), (1e6, 8), (1000003, 4), (7, 8)):
        parts = [shard_packets(packages, r, n) for r in range(n)]
        assert len({p[2] for p in parts}) == 1 and parts[0][2] >= packages
        blocks = sorted((p[1], p[1] + p[0]) for p in parts)
        assert blocks[0][0] == 0 and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    with pytest.raises(ValueError):
        shard_packets(10, 3, 2)


class FakeEngine:
    """records what the host logic asks of the engine"""
    def __init__(self):
        self.calls = []

    def comm_unique_id(self):
        return np.arange(128, dtype=np.uint8)

    def comm_init(self, rank, nranks, uid):
        self.calls.append((rank, nranks, bytes(uid)))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        e = FakeEngine()
        uid = share_unique_id(e, dist)
        npr, offset, total = shard_packets(1001, rank, world)
        # every rank contributes the "packets" of its stream block; the reduction is a plain sum (ncclAllReduce(sum))
        part = torch.zeros(total, dtype=torch.float64); part[offset:offset + npr] = 1.0
        dist.all_reduce(part)
        out.put((rank, e.calls, uid.tolist(), npr, offset, total, part.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_two_ranks_share_the_id_and_partition_the_streams():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = [r[2] for r in res]
    assert ids[0] == ids[1] == list(range(128))
    for rank, calls, *_ in res:
        assert calls == [(rank, 2, bytes(range(128)))]
    assert res[0][3] == res[1][3] == 501 and res[0][5] == 1002
    assert (res[0][4], res[1][4]) == (0, 501)
    assert np.array_equal(res[0][6], np.ones(1002)) and np.array_equal(res[1][6], res[0][6])

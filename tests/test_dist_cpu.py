"""Host-side logic of the multi-GPU path on CPU (gloo, world_size 2): shard ranges, the unique-id broadcast
that bootstraps the NCCL communicator, and the additivity the sample-axis sharding relies on (moments of the
shards sum to the moments of the whole; checked with the oracle standing in for the kernel)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _data
from oracle import oracle as orc
from picard_ica_b200.dist import broadcast_unique_id, shard_range


@pytest.mark.parametrize("t,world", [(10_000_000, 8), (1001, 2), (7, 4), (16, 1), (0, 3), (5, 8)])
def test_shard_range_partitions_samples(t, world):
    prev = 0
    sizes = []
    for r in range(world):
        b, e = shard_range(t, r, world)
        assert b == prev and e >= b
        if r < world - 1:
            assert b % 2 == 0 and e % 2 == 0
        prev = e
        sizes.append(e - b)
    assert prev == t
    assert max(sizes) - min(sizes) <= 3


def test_shard_range_rejects_bad_rank():
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = bytes(range(128)) if rank == 0 else None
        got = broadcast_unique_id(uid, 0)
        assert got == bytes(range(128))
        # sample-sharded moments + allreduce(sum) == moments of the whole matrix
        n, t = 6, 4001
        x = _data.whitened(n, t, seed=3)
        w = _data.orthogonal(n, 5)
        b, e = shard_range(t, rank, world)
        ep = orc.eval_point(x[:, b:e], w, orc.TANH, 1.0, ortho=False, extended=False)
        packed = torch.from_numpy(np.concatenate([ep.gr.ravel(), ep.sd, ep.sq, ep.lrow, ep.hr.ravel()]))
        dist.all_reduce(packed)
        full = orc.eval_point(x, w, orc.TANH, 1.0, ortho=False, extended=False)
        ref = np.concatenate([full.gr.ravel(), full.sd, full.sq, full.lrow, full.hr.ravel()])
        err = float(np.max(np.abs(packed.numpy() - ref)) / np.max(np.abs(ref)))
        out.put((rank, err))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_unique_id_and_moment_additivity():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert [r for r, _ in res] == [0, 1]
    assert all(err <= 1e-12 for _, err in res)

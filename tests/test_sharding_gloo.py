"""Frame sharding over ranks (sister_b200/sharding.py): world_size-2 gloo processes on CPU. The per-rig compute is a
stand-in (there is no GPU here); what is tested is the partition, the ordering and the gather that bench.py and a
multi-GPU caller rely on. On the GPU box the same functions run over NCCL with the Engine as compute_block."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sister_b200.sharding import compute_sharded, frame_shard, shard_counts

H, W = 6, 10


def fake_map(k):
    rng = np.random.default_rng(1000 + k)
    return rng.integers(0, 65536, (H, W), dtype=np.uint16)


def _worker(rank, world, port, n_rigs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def block(idx):
            idx = list(idx)
            if not idx:
                return torch.zeros((0, H, W), dtype=torch.int16)
            return torch.from_numpy(np.stack([fake_map(k) for k in idx]).view(np.int16))

        out = compute_sharded(block, n_rigs, gather_to=0)
        if rank == 0:
            q.put(out.numpy().view(np.uint16).copy())
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_rigs", [5, 4, 1])
def test_two_rank_gather_equals_single_process(n_rigs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rigs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=60)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = np.stack([fake_map(k) for k in range(n_rigs)])
    assert got.shape == want.shape and (got == want).all()


def test_shards_partition_the_batch():
    for n in (0, 1, 7, 8, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [frame_shard(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            c = shard_counts(n, world)
            assert sum(c) == n and max(c) - min(c) <= 1

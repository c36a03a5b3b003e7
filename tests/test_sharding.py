"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard frames, run the (oracle) NMS on their block and
all-gather the compact kept-lane results; the gathered tensor must equal the single-process answer."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from phnet_b200 import sharding, synth  # noqa: E402


def test_shard_range_partitions_exactly():
    for F in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_range(F, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == F
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_pack_unpack():
    keep = torch.tensor([[5, 2, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0]])
    num = torch.tensor([2, 1])
    packed = sharding.pack_kept(keep, num, 4)
    assert packed.shape == (2, 5)
    k, n = sharding.unpack_kept(packed)
    assert torch.equal(k, keep[:, :4]) and torch.equal(n, num)


def _worker(rank, world, port, F, q):
    sys.path.insert(0, ROOT)
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    props, scores = synth.make_frames(F, 64, 36, seed=3)            # same seed everywhere, each rank slices its block
    f0, f1 = sharding.shard_range(F, rank, world)
    keep, num, _ = oracle.nms_batched(props[f0:f1].numpy(), scores[f0:f1].numpy(), None, 50.0, 4, lazy=True, threads=1)
    packed = sharding.pack_kept(torch.from_numpy(keep), torch.from_numpy(num), 4)
    full = sharding.gather_kept(packed, F)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("F", [10, 7])
def test_gloo_world2_gather_matches_single_process(F):
    from oracle import oracle
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, F, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    props, scores = synth.make_frames(F, 64, 36, seed=3)
    keep, num, _ = oracle.nms_batched(props.numpy(), scores.numpy(), None, 50.0, 4, lazy=True)
    want = sharding.pack_kept(torch.from_numpy(keep), torch.from_numpy(num), 4).numpy()
    assert got.shape == want.shape and (got == want).all()

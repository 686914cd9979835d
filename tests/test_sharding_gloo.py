"""CPU: the N>1 host logic (shard by image, gather streams on rank 0) with a world-size-2 gloo group.
The per-rank encoder is replaced by the CPU oracle here — this file tests the sharding, not the kernels."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from tinyimgcodec_b200.sharding import balanced_partition, partition


def test_partition_covers_everything():
    for n in (0, 1, 7, 4096):
        for w in (1, 2, 4, 8):
            parts = partition(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


def test_balanced_partition_ragged():
    px = [1024 * 1024] * 3 + [64 * 64] * 100 + [2048 * 2048]
    parts = balanced_partition(px, 4)
    assert parts[0][0] == 0 and parts[-1][1] == len(px)
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    loads = [sum(px[l:h]) for l, h in parts]
    assert max(loads) <= 0.6 * sum(px)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import oracle_lib as O
    from tests.cases import make_case
    from tinyimgcodec_b200.sharding import compress_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    shapes = [(64, 64), (40, 72), (8, 8), (128, 96), (33, 17), (256, 64), (16, 200)]
    imgs = [make_case({"kind": "synthetic" if i % 2 else "noise", "shape": s, "seed": int(rng.integers(1 << 30))})
            for i, s in enumerate(shapes)]
    out = compress_sharded(imgs, 50, rank, world, lambda ims, q_: [O.compress(im, q_) for im in ims])
    if rank == 0:
        q.put([bytes(o) for o in out] == [O.compress(im, 50) for im in imgs])
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok

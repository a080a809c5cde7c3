"""CPU (gloo, world_size 2): host-side logic of the multi-GPU path -- shard ranges and the image gather."""
import os
import socket

import torch
import torch.multiprocessing as mp

from next3d_b200 import distributed as D


def test_shard_range_covers_everything():
    for total in (1, 7, 8, 240):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env('gloo')
    a, b = D.shard_range(6, r, w)
    imgs = torch.stack([torch.full((3, 4, 4), float(i)) for i in range(a, b)])       # "image" i is filled with its global sample id
    out = D.gather_images(imgs, dst=0)
    t = D.max_over_ranks(10.0 + r, 'cpu')
    if r == 0:
        q.put((out[:, 0, 0, 0].tolist(), t))
    torch.distributed.destroy_process_group()


def test_gather_two_ranks_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ids, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]        # global sample order preserved
    assert tmax == 11.0


def _worker_unequal(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env('gloo')
    spans = [D.shard_range(5, k, w) for k in range(w)]                          # 5 samples over 2 ranks: 3 + 2
    a, b = spans[r]
    imgs = torch.stack([torch.full((3, 2, 2), float(i)) for i in range(a, b)])
    out = D.gather_images(imgs, dst=0, sizes=[e - s for s, e in spans])
    if r == 0:
        q.put(out[:, 0, 0, 0].tolist())
    torch.distributed.destroy_process_group()


def test_gather_unequal_shards_gloo():
    """total % world != 0: shards differ by one image; the gather pads to the largest and trims on the destination."""
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_unequal, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ids = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids == [0.0, 1.0, 2.0, 3.0, 4.0]

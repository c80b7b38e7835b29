"""Multi-GPU host logic on CPU: world size 2 over gloo (the GPU path uses the same functions over NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yololite_b200 import dist as yd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, cap, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = yd.shard_range(n_images, rank, world)
        g = torch.Generator().manual_seed(0)
        counts_all = torch.randint(0, cap + 1, (n_images,), generator=g, dtype=torch.int32)
        boxes_all = torch.rand(n_images, cap, 4, generator=g)
        scores_all = torch.rand(n_images, cap, generator=g)
        classes_all = torch.randint(0, 80, (n_images, cap), generator=g)
        packed = yd.pack_detections(boxes_all[lo:hi], scores_all[lo:hi], classes_all[lo:hi])
        full, cnt = yd.gather_detections(packed, counts_all[lo:hi].contiguous())
        dets = yd.unpack_detections(full, cnt)
        ok = len(dets) == n_images and torch.equal(cnt, counts_all)
        for b, d in enumerate(dets):
            c = int(counts_all[b])
            ok &= torch.equal(d["boxes"], boxes_all[b, :c]) and torch.equal(d["scores"], scores_all[b, :c])
            ok &= torch.equal(d["classes"], classes_all[b, :c])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}


def test_shard_range_covers_batch():
    for n in (1, 7, 64, 65, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [yd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_unpack_overflow_raises():
    import pytest
    packed = torch.zeros(1, 4, 6)
    with pytest.raises(RuntimeError):
        yd.unpack_detections(packed, torch.tensor([4 | yd.OVERFLOW_BIT], dtype=torch.int32))


def _make_packed(boxes, scores, classes, counts, cap):
    """The [b, cap+1, 6] payload as the postprocess kernel writes it (include/yololite_b200.h: yl_postprocess_ex)."""
    n = boxes.shape[0]
    p = torch.zeros(n, cap + 1, 6)
    for b in range(n):
        c = int(counts[b])
        p[b, 0, 0] = c; p[b, 0, 2] = c
        p[b, 1:1 + c, :4] = boxes[b, :c]; p[b, 1:1 + c, 4] = scores[b, :c]; p[b, 1:1 + c, 5] = classes[b, :c].float()
    return p


def _worker_packed(rank, world, port, n_images, cap, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from yololite_b200 import post
        lo, hi = yd.shard_range(n_images, rank, world)
        g = torch.Generator().manual_seed(1)
        counts_all = torch.randint(0, cap + 1, (n_images,), generator=g, dtype=torch.int32)
        boxes_all = torch.rand(n_images, cap, 4, generator=g)
        scores_all = torch.rand(n_images, cap, generator=g)
        classes_all = torch.randint(0, 80, (n_images, cap), generator=g)
        mine = _make_packed(boxes_all[lo:hi], scores_all[lo:hi], classes_all[lo:hi], counts_all[lo:hi], cap)
        full, work = yd.gather_packed(mine, async_op=True)          # ONE collective, overlappable
        work.wait()
        dets = post.unpack(full)
        ok = len(dets) == n_images
        for b, d in enumerate(dets):
            c = int(counts_all[b])
            ok &= torch.equal(d["boxes"], boxes_all[b, :c]) and torch.equal(d["scores"], scores_all[b, :c])
            ok &= torch.equal(d["classes"], classes_all[b, :c])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_single_collective_gather_of_kernel_packed_payload_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, 8, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}


def test_unpack_packed_overflow_raises():
    import pytest
    from yololite_b200 import post
    p = torch.zeros(1, 5, 6)
    p[0, 0, 0] = 4; p[0, 0, 1] = 1; p[0, 0, 2] = 9
    with pytest.raises(RuntimeError):
        post.unpack(p)

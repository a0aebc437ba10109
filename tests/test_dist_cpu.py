"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard the batch, run a stand-in per-frame
"forward" on every rank, gather, and compare with the unsharded result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from millieye_b200.dist import (all_reduce_gradients, all_reduce_sum_, gather_detections, gather_rows, shard_batch, shard_bounds,
                                shard_rows_by_frame)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_detect(frames):
    """Deterministic per-frame stand-in: frame i yields (i % 3) + 1 'detections'."""
    n = frames.shape[0]
    det = torch.zeros(n, 4, 6)
    cnt = torch.zeros(n, dtype=torch.int32)
    for i in range(n):
        k = int(frames[i, 0].item()) % 3 + 1
        cnt[i] = k
        det[i, :k] = frames[i, 0] + torch.arange(k).float()[:, None] / 10
    return det, cnt


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frames = torch.arange(total).float()[:, None].repeat(1, 2)
        mine = shard_batch(frames, world, rank)
        det, cnt = _fake_detect(mine)
        gdet, gcnt = gather_detections(det, cnt)
        gdet, gcnt = gdet.clone(), gcnt.clone()
        # the packed fast path: det and count are views of one flat buffer (ops.NmsBuffers.flat), nothing is copied
        flat = torch.cat((det.reshape(-1), cnt.view(torch.float32)))
        pdet = flat[:det.numel()].view_as(det)
        pcnt = flat[det.numel():].view(torch.int32)
        gdet2, gcnt2 = gather_detections(pdet, pcnt, packed=flat)
        assert torch.equal(gdet2, gdet) and torch.equal(gcnt2, gcnt)
        # variable-length rows with local frame indices -> global
        rows = torch.cat([torch.tensor([[float(i), float(mine[i, 0])]]).repeat(int(cnt[i]), 1) for i in range(len(mine))])
        grows = gather_rows(rows, total, cap=64)
        radar = torch.tensor([[0.0, 1], [1, 2], [total - 1, 3]])
        local_radar = shard_rows_by_frame(radar, total, world, rank)
        # numpy arrays are pickled by value; a torch tensor in an mp.Queue is a shared-memory handle that dies with
        # the worker, and a worker that exits before the parent unpickles gives a sporadic FileNotFoundError
        q.put((rank, gdet.numpy(), gcnt.numpy(), grows.numpy(), local_radar.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_gather_equals_single_process():
    world, total = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=90) for _ in range(world)], key=lambda t: t[0])
    results = [(r[0],) + tuple(torch.from_numpy(a) for a in r[1:]) for r in results]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    frames = torch.arange(total).float()[:, None].repeat(1, 2)
    det, cnt = _fake_detect(frames)
    for rank, gdet, gcnt, grows, local_radar in results:
        assert torch.equal(gdet, det) and torch.equal(gcnt, cnt)          # every rank sees the whole batch
        assert grows.shape[0] == int(cnt.sum())
        assert torch.equal(grows[:, 0], grows[:, 1])                      # local index re-based to the global frame
        lo, hi = shard_bounds(total, world, rank)
        assert all(0 <= v < hi - lo for v in local_radar[:, 0].tolist())
    assert results[0][4][:, 1].tolist() == [1.0, 2.0] and results[1][4][:, 1].tolist() == [3.0]


def _uneven_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total, world, rank)
        rows = torch.tensor([[float(i), float(lo + i)] for i in range(hi - lo) for _ in range(2)])
        grows = gather_rows(rows, total, cap=32)
        err = ""
        try:
            gather_detections(torch.zeros(hi - lo, 2, 3), torch.zeros(hi - lo, dtype=torch.int32))
        except ValueError as e:
            err = str(e)
        q.put((rank, grows.numpy(), err))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_uneven_shards():
    """7 frames on 2 ranks (4 + 3): gather_rows re-bases with shard_bounds, gather_detections refuses unequal shards."""
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_uneven_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=90) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for rank, grows, err in results:
        grows = torch.from_numpy(grows)
        assert grows.shape[0] == 2 * total and torch.equal(grows[:, 0], grows[:, 1])
        assert "pad the batch" in err


def _reduce_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(5, 3)
        unused = torch.nn.Linear(2, 2)                     # never gets a gradient
        x = torch.full((4, 5), float(rank + 1))
        lin(x).sum().backward()
        if rank == 0:
            unused.weight.grad = torch.ones_like(unused.weight)   # only one rank has a gradient for it
        n = all_reduce_gradients(list(lin.parameters()) + list(unused.parameters()))
        summed = lin.weight.grad.clone()
        lin.zero_grad()
        lin(x).sum().backward()
        all_reduce_gradients(list(lin.parameters()), average=True)
        assert torch.allclose(lin.weight.grad * world, summed)
        lin.weight.grad.copy_(summed)
        loss_vec = all_reduce_sum_(torch.arange(10, dtype=torch.float32) * (rank + 1))
        q.put((rank, n, lin.weight.grad.numpy(), unused.weight.grad.numpy(), unused.bias.grad.numpy(), loss_vec.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_bucket_and_loss_reduction():
    """Single-bucket gradient all-reduce (SUM over ranks by default - the stage-3 losses are sums, so this equals the
    whole-batch gradient; missing gradients count as zeros) and the sum-reduction of the stage-3 loss vector, world
    size 2 on gloo."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reduce_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=90) for _ in range(world)], key=lambda t: t[0])
    results = [r[:2] + tuple(torch.from_numpy(a) for a in r[2:]) for r in results]
    for p in procs:
        p.join(timeout=30)
    for rank, n, wgrad, ugrad, ubias, loss_vec in results:
        assert n == 15 + 3 + 4 + 2
        # == the gradient of the unsharded batch (4 rows of ones and 4 rows of twos): 4*1 + 4*2 per weight
        assert torch.allclose(wgrad, torch.full((3, 5), 12.0))
        assert torch.allclose(ugrad, torch.full((2, 2), 1.0)) and torch.equal(ubias, torch.zeros(2))
        assert torch.equal(loss_vec, torch.arange(10, dtype=torch.float32) * 3)
    assert torch.equal(results[0][2], results[1][2])


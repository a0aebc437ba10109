"""dist.PeerGather (copy-engine all-gather over peer memory + stream memory operations) with two processes sharing this
GPU: CUDA IPC, me_peer_copy, me_stream_write_value32 / me_stream_wait_value32 (`-m gpu`).  gloo carries the one-time
exchange of the IPC handles; the per-step traffic is the class's own."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _worker(rank, world, port, steps, numel, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from millieye_b200.dist import PeerGather
        torch.cuda.set_device(0)
        pg = PeerGather(numel, depth=3, device=DEV)
        side = torch.cuda.Stream()
        seen = []
        base = torch.arange(numel, dtype=torch.float32, device=DEV)
        with torch.cuda.stream(side):
            for k in range(steps):
                shard = base * (rank + 1) + 1000.0 * k          # distinct per rank and step
                if rank == 1 and k == 4:
                    torch.cuda._sleep(int(2e8))                  # a slow rank: the fast one has to wait for its flag
                g = pg.gather(shard)
                seen.append(g.clone())                           # same stream: ordered after the arrival waits
        side.synchronize()
        q.put((rank, np.stack([s.cpu().numpy() for s in seen])))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_peer_gather_two_processes():
    import torch.multiprocessing as mp
    steps, numel, world = 9, 4099, 2
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, steps, numel, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    base = np.arange(numel, dtype=np.float32)
    for rank, got in results:
        assert got.shape == (steps, world, numel)
        for k in range(steps):
            for src in range(world):
                np.testing.assert_array_equal(got[k, src], base * (src + 1) + 1000.0 * k, err_msg=f"rank {rank} step {k} src {src}")

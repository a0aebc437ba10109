"""PeerGather timing under torchrun (N ranks, one GPU each): (1) gathers alone, (2) a compute stream next to them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from millieye_b200.dist import PeerGather

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
numel = 557000
pg = PeerGather(numel, 4, dev)
shard = torch.rand(numel, device=dev)
side = torch.cuda.Stream()
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)


def timed(fn, n):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(n)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def gathers(n):
    with torch.cuda.stream(side):
        for _ in range(n):
            pg.gather(shard)


def compute(n):
    for _ in range(n):
        a @ a


def both(n):
    for _ in range(n):
        a @ a
        with torch.cuda.stream(side):
            pg.gather(shard)


def nccl(n):
    out = torch.empty(world * numel, device=dev)
    with torch.cuda.stream(side):
        for _ in range(n):
            dist.all_gather_into_tensor(out, shard)


for name, fn in (("gathers alone", gathers), ("matmul alone", compute), ("matmul + gather per step", both), ("nccl all_gather alone", nccl)):
    fn(5)
    ms = timed(fn, 50)
    if rank == 0:
        print(f"{name:28s} {ms:8.3f} ms / step", flush=True)
dist.destroy_process_group()

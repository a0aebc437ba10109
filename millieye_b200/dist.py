"""Multi-GPU plumbing of the hot path: frames are independent, so a batch is split into
contiguous per-rank shards, every rank runs the same forward on its shard with replicated weights, and
the only exchange is one all-gather of the (capacity-bounded) detections.  One process per GPU,
torch.distributed (NCCL on GPUs; the host logic is tested with gloo on CPU)."""
import torch
import torch.distributed as dist


def shard_bounds(total, world_size, rank):
    """Contiguous split of `total` frames: the first (total % world_size) ranks get one extra frame."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(tensor, world_size, rank):
    lo, hi = shard_bounds(tensor.shape[0], world_size, rank)
    return tensor[lo:hi]


def shard_rows_by_frame(rows, total, world_size, rank):
    """rows (m, k) whose column 0 is a frame index in [0, total): keep this rank's frames, re-based to 0."""
    lo, hi = shard_bounds(total, world_size, rank)
    keep = (rows[:, 0] >= lo) & (rows[:, 0] < hi)
    out = rows[keep].clone()
    out[:, 0] -= lo
    return out


_GATHER_OUT = {}      # (device, dtype, numel, world) -> pre-allocated all_gather output
_SHAPE_CHECKED = set()


def _check_equal_shards(n_local, group):
    """all_gather_into_tensor needs the same shard shape on every rank; checked once per (group, n_local)."""
    key = (id(group), n_local)
    if key in _SHAPE_CHECKED:
        return
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
    every = torch.zeros((world,), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(every, mine, group=group)
    if not bool((every == n_local).all()):
        raise ValueError(f"gather_detections: shards hold {every.tolist()} frames; pad the batch so that every rank "
                         "runs the same number of frames")
    _SHAPE_CHECKED.add(key)


def gather_detections(det, count, group=None, packed=None, out=None):
    """det (n_local, max_det, cols) fp32, count (n_local,) int32 -> the same for the whole batch on every rank, with
    ONE collective: counts travel in the same buffer as the rows (bit-cast to fp32) and the output is a
    pre-allocated tensor (all_gather_into_tensor; no list form, no cat).  `packed`: the flat fp32 buffer det and count
    are views of (ops.NmsBuffers.flat) - without it the two are packed with one copy.  Every rank must hold the same
    n_local (checked once per shape).  The returned tensors are views of `out` (flat fp32, world * (det.numel() +
    n_local) elements, owned by the caller) or, without it, of a cached buffer that the next call overwrites."""
    world = dist.get_world_size(group)
    if world == 1:
        return det, count
    n, max_det, cols = det.shape
    _check_equal_shards(n, group)
    body = n * max_det * cols
    if packed is None or packed.numel() != body + n or packed.data_ptr() != det.data_ptr():
        packed = torch.cat((det.reshape(-1), count.view(torch.float32).reshape(-1)))
    key = (packed.device, packed.numel(), world, id(group))
    if out is None:
        out = _GATHER_OUT.get(key)
    if out is None:
        # flat (concatenation along dim 0): the layout both NCCL and gloo accept for all_gather_into_tensor
        out = _GATHER_OUT[key] = torch.empty((world * packed.numel(),), dtype=torch.float32, device=packed.device)
    dist.all_gather_into_tensor(out, packed.reshape(-1), group=group)
    out2 = out.view(world, packed.numel())
    # one strided copy per tensor (the rows of the ranks are not adjacent in the gathered buffer)
    gdet = out2[:, :body].reshape(world * n, max_det, cols)
    gcount = out2[:, body:].view(torch.int32).reshape(world * n)
    return gdet, gcount


def gather_rows(rows, total_frames, cap, group=None):
    """Variable-length (k, cols) result rows whose column 0 is a local frame index -> all ranks' rows with
    global frame indices, in rank order.  `total_frames` is the size of the whole batch: this rank's first global
    frame is shard_bounds(total_frames, world, rank)[0], also for uneven shards.  Two-phase: counts, then payload
    padded to `cap` rows."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return rows
    k = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    ks = torch.zeros((world,), dtype=torch.int64, device=rows.device)
    dist.all_gather_into_tensor(ks, k, group=group)
    ks = ks.tolist()
    if max(ks) > cap:
        raise ValueError("gather_rows: cap too small for the largest shard result")
    payload = torch.zeros((cap, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    payload[:rows.shape[0]] = rows
    payload[:rows.shape[0], 0] += shard_bounds(total_frames, world, rank)[0]
    bufs = torch.empty((world * payload.numel(),), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(bufs, payload.reshape(-1), group=group)
    bufs = bufs.view(world, cap, rows.shape[1])
    return torch.cat([bufs[r, :ks[r]] for r in range(world)], 0)


def all_reduce_sum_(t, group=None):
    """In-place sum over ranks (no-op for a single process).  The stage-3 losses are sums over proposals
    (reduction='sum', reference my_models.py:296-313,614-633) and the metric entries are counts, so the whole-batch
    values of a sharded step are the element-wise sums of the ranks' me_stage3_loss vectors."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_gradients(params, group=None, average=False):
    """One flat bucket for every gradient of `params` (the stage-3 trainable set is 112 845 fp32 values, SURVEY.md
    F8: latency-bound, so a single all-reduce): flatten, all-reduce, scatter back.  Parameters without a gradient
    (F7: unused heads, or no image proposals on this rank) take part with zeros so that every rank reduces the same
    bucket layout.  Returns the number of reduced elements.

    The gradients are SUMMED by default: the stage-3 losses are sums over proposals (reduction='sum', reference
    my_models.py:296-313,614-633; all_reduce_sum_ adds the ranks' loss vectors the same way), so the sum over ranks of
    the shard gradients is the gradient of the whole-batch step a single process would take.  average=True divides by
    the world size (the DistributedDataParallel convention for mean-reduced losses)."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in params])
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return off


class PeerGather:
    """All-gather of equal-sized flat fp32 shards over NVLink peer memory WITHOUT a kernel.

    The persistent convolution kernels hold every SM (one 227 KB CTA each), so an NCCL all-gather kernel launched next to
    them waits until they leave and then delays the next one: ~90 us per step at any N > 1 (profiles/round2).  Here every
    rank owns `depth` receive buffers of world x numel floats and one arrival flag per source rank, exposes them to its
    peers through CUDA IPC once, and per step k (all on the caller's stream, copy engines and stream memory operations
    only):
        1. waits until every peer has pushed step k - depth + 1 (its reads of the slot about to be overwritten are
           behind that push in its stream order),
        2. copies its shard into slot k % depth of its own buffer and of every peer's (me_peer_copy),
        3. stores k + 1 into a local word and copies that word onto flag[rank] of every peer (stream order = arrival order),
        4. waits until its own flags of all peers are >= k + 1 (me_stream_wait_value32).
    gather() returns a (world, numel) view of the slot; it stays valid for depth - 1 further calls.  Pushes precede
    waits on every rank, so the exchange cannot deadlock; a rank can run at most `depth` steps ahead of the slowest.

    MEASURED (2 x B200, tools/peer_probe.py, 2.2 MB shards): correct, but NOT faster - a blocked cuStreamWaitValue32 is
    re-polled at a coarse interval (0.37 ms per exchange against 0.03 ms for ncclAllGather) and delays work queued on the
    process's other streams (a matmul loop next to it went from 0.67 to 1.29 ms per step).  DetectPipeline therefore keeps
    the NCCL collective as its default and offers this class as gather="peer" only."""

    def __init__(self, numel, depth, device, group=None):
        from torch.multiprocessing.reductions import reduce_tensor
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.numel, self.depth, self.device = int(numel), max(2, int(depth)), torch.device(device)
        self.step = 0
        with torch.cuda.device(self.device):
            self.data = torch.zeros((self.depth, self.world, self.numel), dtype=torch.float32, device=self.device)
            self.flags = torch.zeros((self.world,), dtype=torch.int32, device=self.device)
            self.ticks = torch.zeros((2 * self.depth,), dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
        mine = (reduce_tensor(self.data), reduce_tensor(self.flags))
        every = [None] * self.world
        dist.all_gather_object(every, mine, group=group)
        self.peer_data, self.peer_flags = {}, {}
        for p, (d, f) in enumerate(every):
            if p != self.rank:
                self.peer_data[p] = d[0](*d[1])       # rebuild_cuda_tensor: opens the peer's allocation (cudaIpcOpenMemHandle)
                self.peer_flags[p] = f[0](*f[1])
        from . import _lib
        with torch.cuda.device(self.device):           # direct NVLink access from this GPU to every peer's buffers
            for p, buf in self.peer_data.items():
                _lib.check(_lib.lib().me_peer_enable(int(buf.device.index)), "me_peer_enable")
        dist.barrier(group=group)                      # nobody pushes before every rank has opened every buffer

    def gather(self, shard):
        from . import _lib
        from ._lib import check, ptr, stream_ptr
        L = _lib.lib()
        assert shard.is_cuda and shard.dtype == torch.float32 and shard.is_contiguous() and shard.numel() == self.numel
        k, st = self.step, stream_ptr()
        self.step += 1
        slot = k % self.depth
        nbytes = self.numel * 4
        if k >= self.depth:
            for p in self.peer_flags:
                # step j stores j + 1: the peer has pushed step k - depth + 1, i.e. is done with step k - depth's slot
                check(L.me_stream_wait_value32(self.flags[p:p + 1].data_ptr(), k - self.depth + 2, st), "me_stream_wait_value32")
        check(L.me_peer_copy(self.data[slot, self.rank].data_ptr(), ptr(shard), nbytes, st), "me_peer_copy")
        for p, buf in self.peer_data.items():
            check(L.me_peer_copy(buf[slot, self.rank].data_ptr(), ptr(shard), nbytes, st), "me_peer_copy")
        tick = self.ticks[k % (2 * self.depth):k % (2 * self.depth) + 1]
        check(L.me_stream_write_value32(tick.data_ptr(), k + 1, st), "me_stream_write_value32")
        for p, fl in self.peer_flags.items():
            check(L.me_peer_copy(fl[self.rank:self.rank + 1].data_ptr(), tick.data_ptr(), 4, st), "me_peer_copy")
        for p in self.peer_flags:
            check(L.me_stream_wait_value32(self.flags[p:p + 1].data_ptr(), k + 1, st), "me_stream_wait_value32")
        return self.data[slot]

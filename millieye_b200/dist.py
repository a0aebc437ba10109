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


def gather_detections(det, count, group=None):
    """det (n_local, max_det, cols), count (n_local,) int32 -> the same for the whole batch on every rank.
    Shards must have equal n_local (pad the batch otherwise); one all_gather per tensor."""
    world = dist.get_world_size(group)
    if world == 1:
        return det, count
    dets = [torch.empty_like(det) for _ in range(world)]
    counts = [torch.empty_like(count) for _ in range(world)]
    dist.all_gather(dets, det.contiguous(), group=group)
    dist.all_gather(counts, count.contiguous(), group=group)
    return torch.cat(dets, 0), torch.cat(counts, 0)


def gather_rows(rows, frames_per_rank, cap, group=None):
    """Variable-length (k, cols) result rows whose column 0 is a local frame index -> all ranks' rows with
    global frame indices, in rank order.  Two-phase: counts, then payload padded to `cap` rows."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return rows
    k = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    ks = [torch.zeros_like(k) for _ in range(world)]
    dist.all_gather(ks, k, group=group)
    if int(max(v.item() for v in ks)) > cap:
        raise ValueError("gather_rows: cap too small for the largest shard result")
    payload = torch.zeros((cap, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    payload[:rows.shape[0]] = rows
    payload[:rows.shape[0], 0] += rank * frames_per_rank
    bufs = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(bufs, payload, group=group)
    return torch.cat([b[:int(c.item())] for b, c in zip(bufs, ks)], 0)


def all_reduce_sum_(t, group=None):
    """In-place sum over ranks (no-op for a single process).  The stage-3 losses are sums over proposals
    (reduction='sum', reference my_models.py:296-313,614-633) and the metric entries are counts, so the whole-batch
    values of a sharded step are the element-wise sums of the ranks' me_stage3_loss vectors."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_gradients(params, group=None, average=True):
    """One flat bucket for every gradient of `params` (the stage-3 trainable set is 112 845 fp32 values, SURVEY.md
    F8: latency-bound, so a single all-reduce): flatten, all-reduce, scatter back.  Parameters without a gradient
    (F7: unused heads, or no image proposals on this rank) take part with zeros so that every rank reduces the same
    bucket layout.  Returns the number of reduced elements."""
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

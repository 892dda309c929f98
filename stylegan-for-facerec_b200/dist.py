"""Batch sharding for multi-GPU synthesis: one process per GPU, no data-path collective.

Samples are independent all the way through the generator (no BatchNorm, no cross-sample op,
SURVEY.md section 8e), so N GPUs run N shards of the latents with replicated weights.  The reference's
equivalent is `nn.DataParallel` scatter/gather (coach_restyle_psp.py:134-135).
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the samples rank owns: contiguous, sizes differ by at most one, ragged n allowed."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, world_size: Optional[int] = None, rank: Optional[int] = None) -> torch.Tensor:
    """This rank's slice along dim 0 (a view, no copy)."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(x.shape[0], world_size, rank)
    return x[lo:hi]


def gather_images(local: torch.Tensor, total: int) -> torch.Tensor:
    """Optional: assemble [total, ...] on every rank from ragged per-rank shards (pads to the largest
    shard for the all_gather, then trims).  Not needed for synthesis itself."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)


def synthesize_sharded(generator, styles, total: Optional[int] = None, gather: bool = False, **kwargs):
    """Run `generator(styles, **kwargs)` on this rank's shard of every style tensor."""
    local = [shard_batch(s) for s in styles]
    image, aux = generator(local, **kwargs)
    if gather:
        image = gather_images(image, total if total is not None else styles[0].shape[0])
    return image, aux


# ---- the one exchange step of the fine-tuning path -------------------------------------
# ReStyle trains the ENCODER only; the decoder is frozen and replicated, so the only collective of a training step is the
# average of the encoder's gradients.  The reference gets it implicitly from nn.DataParallel's reduce-add onto GPU 0
# (coach_restyle_psp.py:134-135); here every rank keeps a full replica and the gradients are averaged in place with
# bucketed, asynchronous all-reduces (NCCL over NVLink / NVSwitch on a B200 box, gloo in the CPU tests).
def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


def average_gradients(params, bucket_bytes: int = 64 << 20) -> int:
    """All-reduce (mean) the .grad of `params` across ranks, in place.  Gradients are packed into flat buckets of at most
    `bucket_bytes` per (dtype, device) so that a 535 MB encoder costs a few dozen launches, every bucket's all-reduce is
    issued asynchronously before the first one is waited for, and a parameter without a gradient on this rank contributes
    zeros (the bucket layout must not depend on the rank).  Returns the number of buckets (collectives) issued."""
    params = [p for p in params if p.requires_grad]
    if not dist.is_initialized() or dist.get_world_size() == 1 or not params:
        return 0
    world = dist.get_world_size()
    buckets, cur, cur_bytes, cur_key = [], [], 0, None
    for p in params:
        key = (p.dtype, p.device)
        nbytes = p.numel() * p.element_size()
        if cur and (key != cur_key or cur_bytes + nbytes > bucket_bytes):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += nbytes
        cur_key = key
    if cur:
        buckets.append(cur)
    pending = []
    for group in buckets:
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in group])
        pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, group))
    for work, flat, group in pending:
        work.wait()
        flat.div_(world)
        off = 0
        for p in group:
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
    return len(buckets)

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
    """Run `generator(styles, **kwargs)` on this rank's shard of every style tensor.  Per-sample `noise` maps
    ([B_global, 1, r, r]) are sharded with the latents; maps broadcast over the batch ([1, 1, r, r]) pass through."""
    local = [shard_batch(s) for s in styles]
    n_global = styles[0].shape[0]
    noise = kwargs.get("noise")
    if noise is not None:
        kwargs = dict(kwargs)
        kwargs["noise"] = [n if (n is None or n.shape[0] != n_global or n_global == 1) else shard_batch(n) for n in noise]
    image, aux = generator(local, **kwargs)
    if gather:
        image = gather_images(image, total if total is not None else n_global)
    return image, aux


# ---- the one exchange step of the fine-tuning path -------------------------------------
# ReStyle trains the ENCODER only; the decoder is frozen and replicated, so the only collective of a training step is the
# average of the encoder's gradients.  The reference gets it implicitly from nn.DataParallel's reduce-add onto GPU 0
# (coach_restyle_psp.py:134-135); here every rank keeps a full replica and the gradients are averaged in place with
# bucketed, asynchronous all-reduces (NCCL over NVLink / NVSwitch on a B200 box, gloo in the CPU tests).
def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """Make every rank start from rank `src`'s parameters and buffers.  The broadcast writes through `t.detach()`, which
    shares the tensor's version counter (a write through `.data` does not bump it), and then drops every packed-weight
    cache below `module` (`invalidate_caches()`), so the bf16 engine re-packs from the new values."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.detach(), src)
    for m in module.modules():
        inv = getattr(m, "invalidate_caches", None)
        if callable(inv):
            inv()


def _bucket_layout(params, bucket_bytes):
    """consecutive parameters of one (dtype, device) packed into buckets of at most bucket_bytes"""
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
    return buckets


class GradientAverager:
    """Bucketed gradient all-reduce (mean) that overlaps with the backward pass which produces the gradients.

    * The gradients LIVE in flat buckets: every `p.grad` is a view into its bucket (no concatenation, no copy back), the
      buckets are laid out in REVERSE parameter order, i.e. roughly the order in which autograd finishes them.
    * `arm()` before the last backward of a step (ReStyle accumulates `n_iters_per_batch` = 5 backward passes into the
      same gradients, coach_restyle_psp.py:138-168; only the last one may trigger the exchange): a post-accumulate hook
      per parameter counts its bucket down and issues the bucket's asynchronous all-reduce the moment it is complete,
      while autograd is still working on the earlier layers -- over NCCL the collective runs on its own stream.
    * `finish()` issues whatever was not triggered (parameters that received no gradient in the armed pass), waits, and
      turns sums into means.  Parameters that received a gradient on NO rank during the step get `grad = None` back, so
      an optimizer with weight decay or momentum skips them exactly like the single-process run would.
    """

    def __init__(self, params, bucket_bytes: int = 64 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        # NCCL averages inside the collective (ReduceOp.AVG); gloo only sums
        self._avg = self.world > 1 and dist.get_backend() == "nccl"
        self.buckets = []
        self._index = {}
        for group in _bucket_layout(list(reversed(self.params)), bucket_bytes):
            flat = torch.zeros(sum(p.numel() for p in group), dtype=group[0].dtype, device=group[0].device)
            views, off = [], 0
            for p in group:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            b = {"flat": flat, "params": group, "views": views, "pending": 0, "work": None}
            for k, p in enumerate(group):
                self._index[id(p)] = (len(self.buckets), k)
            self.buckets.append(b)
        self._pos = {id(p): i for i, p in enumerate(self.params)}
        self._touched = torch.zeros(len(self.params), dtype=torch.float32)
        self._armed = False
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.attach()

    def attach(self):
        """(re)point every p.grad at its bucket view; the buckets are zeroed (the start of an optimizer step)"""
        for b in self.buckets:
            b["flat"].zero_()
            b["work"] = None
            for p, v in zip(b["params"], b["views"]):
                p.grad = v
        self._touched.zero_()
        self._armed = False

    zero_grad = attach

    def arm(self):
        self._armed = True
        for b in self.buckets:
            b["pending"] = len(b["params"])

    def _on_grad(self, p):
        self._touched[self._pos[id(p)]] = 1.0
        bi, k = self._index[id(p)]
        b = self.buckets[bi]
        if p.grad is not b["views"][k]:                    # someone replaced .grad (e.g. zero_grad(set_to_none=True)): fold it back
            b["views"][k].copy_(p.grad)
            p.grad = b["views"][k]
        if self._armed:
            b["pending"] -= 1
            if b["pending"] == 0 and self.world > 1 and b["work"] is None:
                b["work"] = dist.all_reduce(b["flat"], op=self._op(), async_op=True)

    def _op(self):
        return dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM

    def finish(self) -> int:
        """-> number of collectives issued for this step"""
        n = 0
        untouched_here = bool((self._touched == 0).any())
        if self.world > 1:
            for b in self.buckets:
                if b["work"] is None:
                    b["work"] = dist.all_reduce(b["flat"], op=self._op(), async_op=True)
            # which parameters received a gradient on ANY rank: every rank takes part in the exchange, only a rank that
            # has an untouched parameter itself reads the answer back (a host synchronisation the common case never pays)
            mask = self._touched.to(self.buckets[0]["flat"].device, non_blocking=True) if self.buckets else self._touched
            mwork = dist.all_reduce(mask, op=dist.ReduceOp.MAX, async_op=True) if self.buckets else None
            for b in self.buckets:
                b["work"].wait()
                if not self._avg:
                    b["flat"].div_(self.world)
                b["work"] = None
                n += 1
            if mwork is not None:
                mwork.wait()
                n += 1
                if untouched_here:
                    self._touched = mask.cpu()
        if untouched_here:
            for p in self.params:
                if self._touched[self._pos[id(p)]] == 0:
                    p.grad = None
        self._armed = False
        return n

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def average_gradients(params, bucket_bytes: int = 64 << 20) -> int:
    """All-reduce (mean) the .grad of `params` across ranks, in place, after the backward passes are done (the blocking
    form; `GradientAverager` is the one that overlaps with backward).  Gradients are packed into flat buckets of at most
    `bucket_bytes` per (dtype, device) with one fused multi-tensor copy each way, every bucket's all-reduce is issued
    asynchronously before the first one is waited for, and a parameter without a gradient on this rank contributes zeros
    (the bucket layout must not depend on the rank) -- unless it has a gradient on NO rank, in which case its `.grad`
    stays None.  Returns the number of gradient buckets (collectives) issued."""
    params = [p for p in params if p.requires_grad]
    if not dist.is_initialized() or dist.get_world_size() == 1 or not params:
        return 0
    world = dist.get_world_size()
    buckets = _bucket_layout(params, bucket_bytes)
    has = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], device=params[0].device)
    mwork = dist.all_reduce(has, op=dist.ReduceOp.MAX, async_op=True)
    pending = []
    for group in buckets:
        flat = torch.zeros(sum(p.numel() for p in group), dtype=group[0].dtype, device=group[0].device)
        views, off = [], 0
        for p in group:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        src = [(v, p.grad) for v, p in zip(views, group) if p.grad is not None]
        if src:
            torch._foreach_copy_([v for v, _ in src], [g for _, g in src])
        pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, views, group))
    mwork.wait()
    has = has.cpu()
    anywhere = {id(p): bool(has[i] > 0) for i, p in enumerate(params)}
    for work, flat, views, group in pending:
        work.wait()
        flat.div_(world)
        for p, v in zip(group, views):
            if not anywhere[id(p)]:
                continue                                   # no rank produced a gradient: leave .grad = None
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
    return len(buckets)

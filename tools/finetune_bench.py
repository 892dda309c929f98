#!/usr/bin/env python
"""BASELINE.json configs[2] as a measured workload: the ReStyle fine-tuning step on 1/2/4/8 B200 (one process per GPU).

    python tools/finetune_bench.py [--batch 8] [--steps 10] [--warmup 3]            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/finetune_bench.py ...                                                # N GPUs, NCCL
    python bench.py --workload finetune ...                                        # the same through bench.py

One step = `perform_train_iteration_on_batch` + the optimizer step of the reference (training/coach_restyle_psp.py:138-168,
178-187): 5 refinement iterations, each `codes = encoder(cat[x, y_hat]) + latent` (models/psp.py:84-91), frozen
StyleGAN2-256 decoder forward on the sg2_b200 autograd path (tensor-core route), `face_pool`, bilinear resize to 112
(coach :156), MSE against the target, `loss.backward()` through the decoder into the encoder; then ONE exchange step --
the mean of the encoder's gradients over the ranks (the reference: nn.DataParallel's reduce-add, coach :134-135; here
`dist.GradientAverager`: buckets issued over NCCL from gradient hooks DURING the fifth backward) -- and Adam.

The encoder itself (an IR-SE-50 with 14 map2style heads, 535 MB of fp32 parameters) is out of this repository's scope
(SURVEY.md section 2); a synthetic stand-in with the same parameter volume, the same input ([B, 6, 112, 112]) and output
([B, 14, 512] residual codes) and a layer-by-layer backward keeps the gradient exchange and its overlap honest.

Prints ONE JSON line: images/s (B x ranks / step time, CUDA events, max over ranks), the decoder's share, the exposed
(non-overlapped) time of the exchange, the same step with the blocking exchange after the backward, and the bus
bandwidth of the gradient all-reduce on its own."""
import argparse
import importlib
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class SyntheticEncoder(torch.nn.Module):
    """stand-in for BackboneEncoder(50, 'ir_se', 14, input_nc=6): [B, 6, 112, 112] -> [B, n_styles, 512], ~134 M parameters"""

    def __init__(self, n_styles=14, width=4096, depth=6):
        super().__init__()
        self.pool = torch.nn.AvgPool2d(7)                                  # 112 -> 16
        dims = [6 * 16 * 16] + [width] * (depth + 1)
        self.body = torch.nn.ModuleList(torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.heads = torch.nn.Linear(width, n_styles * 512)
        self.n_styles = n_styles
        with torch.no_grad():
            self.heads.weight.mul_(0.05)

    def forward(self, x):
        h = self.pool(x).flatten(1)
        for l in self.body:
            h = torch.nn.functional.leaky_relu(l(h), 0.2)
        return self.heads(h).view(-1, self.n_styles, 512)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step (weak scaling)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--iters-per-batch", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "exact"])
    ap.add_argument("--bucket-mb", type=int, default=64)
    ap.add_argument("--gpus", type=int, default=None)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sg2 = importlib.import_module("stylegan-for-facerec_b200")
    D = importlib.import_module("stylegan-for-facerec_b200.dist")
    psp_io = importlib.import_module("stylegan-for-facerec_b200.psp_io")

    torch.manual_seed(0)
    G = sg2.Generator(a.size, 512, 8).to(dev).eval()
    G.precision = a.precision
    for p in G.parameters():
        p.requires_grad_(False)
    torch.manual_seed(100 + rank)                              # replicas differ until the broadcast
    enc = SyntheticEncoder(G.n_latent).to(dev)
    D.broadcast_parameters(enc, 0)
    params = list(enc.parameters())
    grad_bytes = sum(p.numel() * 4 for p in params)
    opt = torch.optim.Adam(params, lr=1e-4, fused=True)
    B, K, W, T = a.batch, a.steps, a.warmup, a.iters_per_batch
    g = torch.Generator().manual_seed(7 + rank)
    x = (torch.rand(B, 3, 112, 112, generator=g) * 2 - 1).to(dev)
    y = (torch.rand(B, 3, 112, 112, generator=g) * 2 - 1).to(dev)
    with torch.no_grad():
        latent_avg = G.mean_latent(4096).detach()                                      # psp.py:93-96
        avg_img, _ = G([latent_avg.unsqueeze(1).repeat(1, G.n_latent, 1)], input_is_latent=True, randomize_noise=False)
        avg_img = psp_io.resize_bilinear(avg_img[..., 35:35 + 188, 30:30 + 188].contiguous(), 112)   # coach :81-82
    latent0 = latent_avg.unsqueeze(1).repeat(B, G.n_latent, 1)
    averager = D.GradientAverager(params, bucket_bytes=a.bucket_mb << 20)
    t_dec = [0.0]
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(overlap, time_decoder=False):
        averager.zero_grad()
        y_hat, latent = avg_img.repeat(B, 1, 1, 1), latent0
        marks = []
        for it in range(T):
            x_in = torch.cat([x, y_hat.detach()], 1)
            codes = enc(x_in) + latent.detach()
            if time_decoder:
                e0 = ev(); e0.record()
            img, _ = G([codes], input_is_latent=True, randomize_noise=False)
            y_hat = psp_io.resize_bilinear(psp_io.face_pool(img, (256, 256)), 112)
            loss = torch.nn.functional.mse_loss(y_hat, y)
            if overlap and it == T - 1:
                averager.arm()
            loss.backward()
            if time_decoder:
                e1 = ev(); e1.record(); marks.append((e0, e1))
            latent = codes
        e2 = ev(); e2.record()
        n = averager.finish()
        e3 = ev(); e3.record()
        opt.step()
        return loss, n, marks, (e2, e3)

    def timed(overlap):
        for _ in range(W):
            step(overlap)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        tails = []
        e0.record()
        for _ in range(K):
            loss, n, _, tail = step(overlap)
            tails.append(tail)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / K, sum(s.elapsed_time(e) for s, e in tails) / K], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), n, float(loss.detach())

    ms_ov, tail_ov, n_coll, loss_ov = timed(True)
    ms_blk, tail_blk, _, _ = timed(False)
    # decoder share (forward + backward incl. the encoder's backward of that iteration, CUDA events, no overlap)
    _, _, marks, _ = step(False, time_decoder=True)
    torch.cuda.synchronize()
    dec_ms = sum(s.elapsed_time(e) for s, e in marks)
    # the gradient all-reduce on its own: bus bandwidth over NVLink / NVSwitch
    ar = None
    if world > 1:
        flats = [b["flat"] for b in averager.buckets]
        for _ in range(3):
            for f in flats:
                dist.all_reduce(f)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = ev(), ev()
        e0.record()
        reps = 10
        for _ in range(reps):
            works = [dist.all_reduce(f, async_op=True) for f in flats]
            for w_ in works:
                w_.wait()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        ar = {"standalone_ms": round(ms, 3), "algbw_gbs": round(grad_bytes / ms / 1e6, 1),
              "busbw_gbs": round(2 * (world - 1) / world * grad_bytes / ms / 1e6, 1),
              "reference_busbw_gbs": 725.0, "note": "all buckets in flight at once; bus bandwidth = 2(N-1)/N * bytes / time "
              "(B200_PROFILING.md quotes 725 GB/s for an 8-rank all-reduce at 1 GiB)"}
    if rank == 0:
        line = {"metric": "ReStyle fine-tuning images/sec (5 refinement iterations, fwd+bwd through the StyleGAN2-256 decoder)",
                "value": round(B * world / ms_ov * 1e3, 2), "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms_ov, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if a.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[2]: {T} x (synthetic encoder -> StyleGAN2-{a.size} decoder fwd+bwd, face_pool, "
                                       f"bilinear 112, MSE) + gradient all-reduce + Adam, batch {B} per GPU",
                           "batch_per_gpu": B, "global_batch": B * world, "iters_per_batch": T,
                           "encoder": f"synthetic MLP stand-in, {grad_bytes / 1e6:.0f} MB fp32 gradients (IR-SE-50 + 14 heads: 535 MB)",
                           "decoder": f"sg2_b200 Generator({a.size}), frozen, precision={a.precision}",
                           "parallelism": f"data-parallel x{world}, one NCCL all-reduce (mean) of the encoder gradients per step, "
                                          f"{a.bucket_mb} MiB buckets issued from gradient hooks during the last backward"},
                "decoder_ms_per_step": round(dec_ms, 3), "decoder_passes_per_s": round(B * world * T / ms_ov * 1e3, 1),
                "exchange": {"gradient_bytes": grad_bytes, "collectives_per_step": n_coll,
                             "exposed_ms_overlapped": round(tail_ov, 3), "exposed_ms_blocking": round(tail_blk, 3),
                             "ms_per_step_blocking": round(ms_blk, 3), "allreduce": ar},
                "loss": round(loss_ov, 5)}
        print(json.dumps(line), flush=True)
    averager.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

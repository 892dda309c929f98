#!/usr/bin/env python
"""bench.py -- StyleGAN2 synthesis throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's CPU path, host cores)

A "step" = one Generator.forward over one batch of synthetic latents (z ~ N(0,1), mapping network
included, randomize_noise=False noise buffers), i.e. BASELINE.json configs[1]:
"StyleGAN2 256x256 synthesis forward bf16, batch 64, 1xB200".  Batch is per GPU (weak scaling, the
path shards by sample, no collective in the data path).

  value : images/s with the latents already resident in HBM (CUDA events around K steps, max over ranks)
  e2e   : images/s through the public module API with HOST buffers: every step copies its latents
          from pinned host memory and reads the produced images back into pinned host memory
  roofline : the tcgen05 implicit-GEMM kernel (ModulatedConv2d): algorithmic FLOPs of all its launches
          in a step / their summed CUDA-event durations, vs the measured dense bf16 peak
  cpu_baseline : the reference's algorithm on the host cores (oracle port = the reference's model.py
          math on torch-CPU ops, see oracle/sg2_oracle.py), bounded sample
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE_DEFAULT, BATCH_DEFAULT, N_MLP, STYLE_DIM = 256, 64, 8, 512


def workload_name(size, batch):
    """the one workload both arms report (BASELINE.json configs[1] at the defaults)"""
    return (f"StyleGAN2 config-f {size}x{size} synthesis forward (mapping + synthesis), random-init weights, "
            f"batch {batch} per GPU")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "which": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "which": "fallback"}


def ncu_traffic(size, batch):
    """DRAM bytes (read + write) per step of the dominant kernel, from the committed `ncu --set full` capture of one
    forward at this size / batch (profiles/ncu_full_*_summary.json, written by tools/ncu_summary.py): -> (bytes, file)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"ncu_full_r*_{size}_b{batch}_summary.json")), reverse=True):
        try:
            rows = json.load(open(path))
        except Exception:
            continue
        gemm = [r for r in rows if "modconv_gemm" in r.get("kernel", "")]
        if gemm:
            tot = sum(r["dram__bytes_read.sum"]["value"] + r["dram__bytes_write.sum"]["value"] for r in gemm)
            return tot, os.path.relpath(path, ROOT), len(gemm)
    return None, None, 0


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower() == "active":
                        reasons.add(name)
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_generator(sg2, size, device, precision):
    """SURVEY.md section 8d synthetic init: default init, then noise/bias parameters ~ N(0, 0.1^2)."""
    torch.manual_seed(0)
    G = sg2.Generator(size, STYLE_DIM, N_MLP)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in G.named_parameters():
            if name.endswith("noise.weight") or name.endswith("activate.bias") or (name.endswith(".bias") and "to_rgb" in name):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    G = G.to(device).eval()
    G.precision = precision
    return G


def cpu_baseline(size, batch, iters, warm):
    """oracle port (the reference's model.py math on torch-CPU ops), all host threads."""
    from oracle import sg2_oracle as O
    torch.set_num_threads(os.cpu_count())
    sd = O.init_state_dict(size, STYLE_DIM, N_MLP)
    z = O.named_randn("bench:cpu:z", (batch, STYLE_DIM), 0)
    ts = []
    with torch.no_grad():
        for i in range(warm + iters):
            t0 = time.perf_counter()
            O.generator_forward(sd, size, [z], n_mlp=N_MLP, randomize_noise=False)
            if i >= warm:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    return batch / ts[len(ts) // 2], ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_b = 4          # BASELINE.json configs[0]: the reference's own CPU-runnable case, batch 4
    ips, ts = cpu_baseline(args.size, sample_b, args.steps, args.warmup)
    line = {"metric": f"StyleGAN2-{args.size} images/sec", "value": round(ips, 4), "unit": "images/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * ts[len(ts) // 2], 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.size, args.batch),
                       "arm": "the reference's model.py math on torch-CPU ops (oracle port), all host threads",
                       "batch_per_gpu": args.batch, "batch_per_step": sample_b, "precision": "fp32"},
            "cpu_baseline": {"value": round(ips, 4), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"each step = one forward of batch {sample_b} (bounded sample of the batch-{args.batch} "
                                       f"GPU step), fp32, randomize_noise=False"},
            "e2e": {"value": round(ips, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=SIZE_DEFAULT)
    ap.add_argument("--batch", type=int, default=BATCH_DEFAULT, help="images per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-iters", type=int, default=5)
    ap.add_argument("--profile-out", default=None, help="write the per-kernel CUDA-event table here (json)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sg2 = importlib.import_module("stylegan-for-facerec_b200")
    peaks = load_peaks()
    B, K, W = args.batch, args.steps, args.warmup

    G = make_generator(sg2, args.size, dev, args.precision)
    gen = torch.Generator().manual_seed(1234 + rank)          # every rank gets its own shard of latents
    z_host = torch.randn(K + W, B, STYLE_DIM, generator=gen).pin_memory()
    z_dev = z_host.to(dev)

    def step(z):
        with torch.no_grad():
            return G([z], randomize_noise=False)[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: latents resident in HBM ---------------------------------------------------------
    for i in range(W):
        step(z_dev[i])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = sg2._lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        img = step(z_dev[W + i])
    e1.record()
    barrier()
    launches = sg2._lib.launch_count() - n0
    elapsed = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    t_dev = float(elapsed.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers, H2D of latents and D2H of images inside the timed region -----------
    # The images leave the device the way the reference's own pipelines consume them: as uint8 (tensor2im,
    # restyle-encoder/utils/common.py:5-11), converted on the device by the repo's public `images_to_uint8`, so a
    # quarter of the bytes crosses PCIe.  The fp32 variant (every byte of the module's output) is timed too and reported
    # as e2e.fp32_images: it is bound by the host link once several ranks copy at the same time.
    psp_io = importlib.import_module("stylegan-for-facerec_b200.psp_io")
    copy_stream = torch.cuda.Stream(device=dev)

    def make_e2e(as_uint8):
        out_host = [torch.empty(B, 3, args.size, args.size, dtype=torch.uint8 if as_uint8 else torch.float32).pin_memory()
                    for _ in range(2)]

        def loop(n, base):
            for i in range(n):
                z = z_host[base + i].to(dev, non_blocking=True)          # H2D from pinned memory
                im = step(z)
                if as_uint8:
                    im = psp_io.images_to_uint8(im)
                ready = torch.cuda.Event()
                ready.record()
                with torch.cuda.stream(copy_stream):                      # D2H overlaps the next step's compute
                    copy_stream.wait_event(ready)
                    out_host[i & 1].copy_(im, non_blocking=True)
                    im.record_stream(copy_stream)
            copy_stream.synchronize()
        return loop

    def time_e2e(as_uint8):
        loop = make_e2e(as_uint8)
        loop(W, 0)
        barrier()
        t0 = time.perf_counter()
        loop(K, W)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_e2e = time_e2e(True)
    t_e2e_f32 = time_e2e(False)

    # ---- roofline of the dominant kernel: per-launch CUDA events inside the engine --------------
    roofline, table = None, None
    if rank == 0 and args.precision == "bf16":
        eng = G.engine()
        desc = [l.split() for l in eng.describe().strip().splitlines()]
        n_k = len(desc)
        import ctypes as C
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_k + 1)]
        for e in evs:
            e.record()                                                # materialise the handles
        torch.cuda.synchronize()
        arr = (C.c_void_p * len(evs))(*[e.cuda_event for e in evs])
        acc = [0.0] * n_k
        reps = min(K, 10)
        w = G.style(z_dev[0])
        lat = w.unsqueeze(1).repeat(1, G.n_latent, 1)
        noise = [getattr(G.noises, f"noise_{i}") for i in range(G.num_layers)]
        for _ in range(reps):
            sg2._lib.check(eng.lib.sg2_synth_set_profile_events(eng.plan, arr, len(evs)))
            eng.synthesize(lat, noise, graph=False)
            torch.cuda.synchronize()
            for k in range(n_k):
                acc[k] += evs[k].elapsed_time(evs[k + 1]) * 1e-3
        sg2._lib.check(eng.lib.sg2_synth_set_profile_events(eng.plan, None, 0))
        table = []
        for k, row in enumerate(desc):
            kv = dict(f.split("=") for f in row[3:])
            table.append({"k": k, "kind": row[1], "what": row[2], "s": acc[k] / reps, "flops": float(kv["flops"]) * B,
                          "bytes": float(kv["bytes"]) * B, "tiles": int(kv["tiles"]), "block_n": int(kv["block_n"])})
        gemm = [r for r in table if r["kind"] == "gemm"]
        fl, tt = sum(r["flops"] for r in gemm), sum(r["s"] for r in gemm)
        total = sum(r["s"] for r in table)
        peak = peaks["bf16_tflops_sustained"]
        traffic, traffic_src, traffic_n = ncu_traffic(args.size, B)
        roofline = {"bound": "tensor", "kernel": "modconv_gemm_kernel (tcgen05 implicit GEMM, all launches of one step)",
                    "achieved": round(fl / tt / 1e12, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(fl / tt / 1e12 / peak, 4),
                    "traffic": traffic, "traffic_unit": "DRAM bytes per step (read + write), all launches of the kernel",
                    "traffic_source": traffic_src if traffic_n == len(gemm) else
                    (f"{traffic_src}: {traffic_n} launches captured, {len(gemm)} in the plan" if traffic_src else None),
                    "algorithmic_bytes_per_step": sum(r["bytes"] for r in gemm),
                    "peak_source": f"{peaks['which']} bf16_tflops_sustained (kernel timed inside a long step)",
                    "share_of_step": round(tt / total, 4), "launches_per_step": len(gemm),
                    "algorithmic_flops_per_step": fl}
        fir = [r for r in table if r["kind"] == "upfir"]
        if fir:
            by, ft = sum(r["bytes"] for r in fir), sum(r["s"] for r in fir)
            roofline["upfir"] = {"bound": "hbm", "achieved": round(by / ft / 1e9, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": round(by / ft / 1e9 / peaks["hbm_gbs"], 4), "share_of_step": round(ft / total, 4)}
        if args.profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
            json.dump({"batch": B, "size": args.size, "reps": reps, "kernels": table}, open(args.profile_out, "w"), indent=1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ips, ts = cpu_baseline(args.size, 4, args.cpu_iters, 1)
        cpu = {"value": round(ips, 3), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"BASELINE configs[0]: {args.size}x{args.size} forward, batch 4, fp32, 1 warm-up + {args.cpu_iters} timed "
                         f"iterations (median {1e3 * ts[len(ts) // 2]:.0f} ms/iter)"}

    n_img = world * B * K
    act_gb = 137.5e6 * B / 1e9 if args.size == 256 else None
    line = {
        "metric": f"StyleGAN2-{args.size} images/sec", "value": round(n_img / t_dev, 2), "unit": "images/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(1e3 * t_dev / K, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.size, B),
                   "batch_per_gpu": B, "global_batch": B * world, "precision": args.precision,
                   "l2": f"no explicit flush: each step streams >= {act_gb:.1f} GB of activations (>> 126 MB L2)" if act_gb else
                         "no explicit flush: per-step activation traffic >> 126 MB L2",
                   "parallelism": f"batch-sharded x{world}, no collective in the data path"},
        "e2e": {"value": round(n_img / t_e2e, 2), "unit": "images/s", "h2d_bytes_per_step": B * STYLE_DIM * 4,
                "d2h_bytes_per_step": B * 3 * args.size * args.size,
                "note": "pinned host latents -> Generator.forward -> images_to_uint8 (tensor2im on the device) -> uint8 images "
                        "copied back to pinned host memory (copy stream overlaps the next step)",
                "fp32_images": {"value": round(n_img / t_e2e_f32, 2), "d2h_bytes_per_step": B * 3 * args.size * args.size * 4,
                                "note": "same loop copying the module's fp32 output instead of uint8"}},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if roofline:
        line["roofline"] = roofline
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- StyleGAN2 synthesis throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's CPU path, host cores)
    python bench.py --workload finetune ...                  (BASELINE configs[2]: tools/finetune_bench.py)

A "step" = one Generator.forward over one batch of synthetic latents (z ~ N(0,1), mapping network
included, randomize_noise=False noise buffers).  The headline is BASELINE.json configs[1]:
"StyleGAN2 256x256 synthesis forward bf16, batch 64, 1xB200"; batch is per GPU (weak scaling, the
path shards by sample, no collective in the data path).  The same run also measures configs[3]
(1024^2, batch 32 per GPU) into `configs["1024_b32"]` and, on rank 0 at N = 1, the reference's own GPU
formulation (cuDNN grouped convolution + the reference's CUDA extensions) into `gpu_reference`.

  value : images/s with the latents already resident in HBM (CUDA events around K steps, max over ranks)
  e2e   : images/s through the public module API with HOST buffers: every step copies its latents
          from pinned host memory and reads the produced images back into pinned host memory
  roofline : the tcgen05 implicit-GEMM kernel (ModulatedConv2d): algorithmic FLOPs of all its launches
          in a step / their summed CUDA-event durations, vs the measured dense bf16 peak
  parity : one sample of the last timed batch recomputed on the exact fp32 path (pinned to the reference's
          golden vectors by tests/) -- the timed engine's image must be finite and within tolerance
  ranks : per-rank min / median / max step time (CUDA events around every step)
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
PARITY_REL_MAX = 1.6e-2          # same bound as tests/test_bench_configs_gpu.py (<= 2x the measured error)


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
        gemm = [r for r in rows if "modconv_" in r.get("kernel", "")]       # modconv_gemm / _gemm2 / _gemm2_poly4 / _dxs
        if gemm:
            tot = sum(r["dram__bytes_read.sum"]["value"] + r["dram__bytes_write.sum"]["value"] for r in gemm)
            return tot, os.path.relpath(path, ROOT), len(gemm)
    return None, None, 0


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    NVML_REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                    ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"),
                    ("hw_power_brake_slowdown", "nvmlClocksEventReasonHwPowerBrakeSlowdown"), ("gpu_idle", "nvmlClocksEventReasonGpuIdle"),
                    ("applications_clocks_setting", "nvmlClocksEventReasonApplicationsClocksSetting"), ("sync_boost", "nvmlClocksEventReasonSyncBoost"))

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.extra, self.source, self._stop = set(), "nvidia-smi -lms 50", False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:                       # the CUDA ordinal is not the NVML index under CUDA_VISIBLE_DEVICES: go by PCI address
            pr = torch.cuda.get_device_properties(self.index)
            return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _nvml_loop(self, nv, h):
        """the same quantities as the nvidia-smi query, read through NVML every ~2 ms: a 0.1 s timed region gets tens of
        samples instead of two or three, and every active clock-event reason is seen, not only the ones sampled by chance"""
        mx = str(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        reasons = [(name, getattr(nv, attr)) for name, attr in self.NVML_REASONS if hasattr(nv, attr)]
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                break
            act = {name for name, bit in reasons if mask & bit}
            flag = lambda n: "Active" if n in act else "Not Active"
            self.rows.append((time.perf_counter(), [str(self.index), str(sm), mx, f"{pw:.2f}", hex(mask), flag("hw_slowdown"),
                                                    flag("hw_thermal_slowdown"), flag("sw_thermal_slowdown"), flag("sw_power_cap"),
                                                    ",".join(sorted(act - {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"}))]))
            time.sleep(0.002)

    def start(self):
        if os.environ.get("SG2_BENCH_CLOCKS", "nvml") == "nvml":
            try:
                nv, h = self._nvml_handle()
                nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                self.proc, self.source = "nvml", "NVML, 2 ms period"
                threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True).start()
                return self
            except Exception:
                self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=3.0):
        """nvidia-smi needs a few hundred ms to print its first row; the caller keeps the GPU busy meanwhile"""
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        return time.perf_counter()

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        if self.proc == "nvml":
            time.sleep(0.01)
            self._stop = True
        else:
            time.sleep(0.12)
            self.proc.terminate()
        return self.summary(t_begin, t_end)

    def summary(self, t_begin=None, t_end=None):
        """clock record of the samples taken in [t_begin, t_end] (perf_counter times); the sampler keeps running"""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        allrows = list(self.rows)
        rows = [r for t, r in allrows if len(r) >= 9 and (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end + 0.06)]
        if not rows:                                   # region shorter than one sampling period: nearest rows
            rows = [r for t, r in allrows if len(r) >= 9 and (t_end is None or t <= t_end + 0.2)][-2:]
        sm = sorted(int(r[1]) for r in rows if r[1].isdigit())
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower() == "active":
                    reasons.add(name)
            if len(r) > 9 and r[9]:                    # NVML sampler: the other clock-event reasons that were active
                reasons.update(r[9].split(","))
        mx = [int(r[2]) for r in rows if r[2].isdigit()]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": sm[0] if sm else None, "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": self.source}


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


def gpu_reference(G, size, batch, dev, iters=5, warm=3):
    """The reference's own GPU formulation on this GPU (`model.py:232-273` -- per-sample weights,
    `F.conv2d(groups=B)` / `F.conv_transpose2d` on cuDNN, fp32) with the reference's own CUDA extensions for
    `upfirdn2d` / `fused_leaky_relu` when oracle/_ref holds them (compiled from /root/reference in the authoring
    container), else the oracle's torch restatement of both ops.  A reported baseline: checker code, never the product."""
    import importlib.util
    from oracle import sg2_oracle as O
    sd = {k: v.detach().float().to(dev) for k, v in G.state_dict().items()}
    z = torch.randn(batch, STYLE_DIM, device=dev)
    ext, saved = {}, (O.upfirdn2d, O.fused_leaky_relu)
    for name in ("fused", "upfirdn2d"):
        path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
        if os.path.exists(path):
            try:
                spec = importlib.util.spec_from_file_location(name, path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                ext[name] = mod
            except Exception:
                pass
    kind = "port: oracle restatement of model.py on CUDA tensors (cuDNN groups=B convolutions, torch ops for upfirdn2d / lrelu)"
    if len(ext) == 2:
        kind = ("reference formulation: model.py math (cuDNN groups=B convolutions) + the reference's own CUDA extensions "
                "(oracle/_ref/{fused,upfirdn2d}.so, sm_100a build of op/*.cu)")

        def up_ext(x, kernel, up=1, down=1, pad=(0, 0)):
            b, c, h, w = x.shape
            out = ext["upfirdn2d"].upfirdn2d(x.reshape(-1, h, w, 1), kernel.to(x.device, torch.float32), up, up, down, down,
                                             pad[0], pad[1], pad[0], pad[1])
            return out.view(b, c, out.shape[1], out.shape[2])

        def lrelu_ext(x, bias, negative_slope=0.2, scale=2 ** 0.5):
            return ext["fused"].fused_bias_act(x.contiguous(), bias, x.new_empty(0), 3, 0, negative_slope, scale)
        O.upfirdn2d, O.fused_leaky_relu = up_ext, lrelu_ext
    res = {"kind": kind, "batch": batch, "size": size, "dtype": "f32", "unit": "images/s"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.benchmark = True
        img_ref = None
        for key, tf32 in (("tf32_off", False), ("tf32_on", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(warm):
                    img, _ = O.generator_forward(sd, size, [z], n_mlp=N_MLP, randomize_noise=False)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters):
                    img, _ = O.generator_forward(sd, size, [z], n_mlp=N_MLP, randomize_noise=False)
                b.record()
                torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            res[key] = {"value": round(batch / ms * 1e3, 2), "ms_per_step": round(ms, 3)}
            if not tf32:
                img_ref = img
        # the engine against this GPU reference on the same latents (true-fp32 cuDNN)
        with torch.no_grad():
            mine, _ = G([z], randomize_noise=False)
        res["engine_vs_this_reference_rel_max"] = float(((mine.float() - img_ref).abs().max() / img_ref.abs().max()).item())
        res["iters"] = iters
    finally:
        O.upfirdn2d, O.fused_leaky_relu = saved
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    return res


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


class Ctx:
    """process-wide state of one bench run"""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.sg2 = importlib.import_module("stylegan-for-facerec_b200")
        self.psp_io = importlib.import_module("stylegan-for-facerec_b200.psp_io")
        self.peaks = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, seconds):
        t = torch.tensor([seconds], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, values):
        """[world, len(values)] on every rank"""
        t = torch.tensor(values, device=self.dev, dtype=torch.float64)
        if self.world == 1:
            return t[None].cpu()
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return torch.stack(out).cpu()


def kernel_table(ctx, G, z, B, reps):
    """per-launch CUDA events inside the engine, taken in a steady-state loop (no host synchronisation between the
    repetitions, so clocks and power state are those of the timed region, not of an idle GPU waking up)"""
    import ctypes as C
    sg2 = ctx.sg2
    eng = G.engine()
    desc = [l.split() for l in eng.describe().strip().splitlines()]
    n_k = len(desc)
    sets = []
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_k + 1)]
        for e in evs:
            e.record()                                                # materialise the handles
        sets.append((evs, (C.c_void_p * len(evs))(*[e.cuda_event for e in evs])))
    torch.cuda.synchronize()
    w = G.style(z)
    lat = w.unsqueeze(1).repeat(1, G.n_latent, 1)
    noise = [getattr(G.noises, f"noise_{i}") for i in range(G.num_layers)]
    for _ in range(3):
        eng.synthesize(lat, noise, graph=False)
    for evs, arr in sets:
        sg2._lib.check(eng.lib.sg2_synth_set_profile_events(eng.plan, arr, len(evs)))
        eng.synthesize(lat, noise, graph=False)
    torch.cuda.synchronize()
    sg2._lib.check(eng.lib.sg2_synth_set_profile_events(eng.plan, None, 0))
    acc = [0.0] * n_k
    for evs, _ in sets:
        for k in range(n_k):
            acc[k] += evs[k].elapsed_time(evs[k + 1]) * 1e-3
    table = []
    for k, row in enumerate(desc):
        kv = dict(f.split("=") for f in row[3:])
        table.append({"k": k, "kind": row[1], "what": row[2], "s": acc[k] / reps, "flops": float(kv["flops"]) * B,
                      "bytes": float(kv["bytes"]) * B, "tiles": int(kv["tiles"]), "block_n": int(kv["block_n"])})
    return table


def roofline_of(ctx, table, size, B):
    peaks = ctx.peaks
    gemm = [r for r in table if r["kind"] == "gemm"]
    fl, tt = sum(r["flops"] for r in gemm), sum(r["s"] for r in gemm)
    total = sum(r["s"] for r in table)
    peak = peaks["bf16_tflops_sustained"]
    traffic, traffic_src, traffic_n = ncu_traffic(size, B)
    by = sum(r["bytes"] for r in gemm)
    roofline = {"bound": "tensor", "kernel": "modconv_gemm* / modconv_dxs (tcgen05 implicit GEMM family, all launches of one step)",
                "achieved": round(fl / tt / 1e12, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(fl / tt / 1e12 / peak, 4),
                "traffic": traffic, "traffic_unit": "DRAM bytes per step (read + write), all launches of the kernel",
                "traffic_source": traffic_src if traffic_n == len(gemm) else
                (f"{traffic_src}: {traffic_n} launches captured, {len(gemm)} in the plan" if traffic_src else None),
                "algorithmic_bytes_per_step": by,
                "hbm_frac": round(by / tt / 1e9 / peaks["hbm_gbs"], 4),
                "peak_source": f"{peaks['which']} bf16_tflops_sustained (kernel timed inside a long step)",
                "share_of_step": round(tt / total, 4), "launches_per_step": len(gemm),
                "algorithmic_flops_per_step": fl, "kernel_sum_ms": round(1e3 * total, 3)}
    fir = [r for r in table if r["kind"] == "upfir"]
    if fir:
        fby, ft = sum(r["bytes"] for r in fir), sum(r["s"] for r in fir)
        roofline["upfir"] = {"bound": "hbm", "achieved": round(fby / ft / 1e9, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": round(fby / ft / 1e9 / peaks["hbm_gbs"], 4), "share_of_step": round(ft / total, 4)}
    return roofline


def make_latents(n_batches, B, rank, size):
    """[n_batches, B, STYLE_DIM] synthetic latents of one rank: every rank gets its own shard ...
    ... except sample 0 of every batch, which is the same on every rank: the parity check recomputes sample 0 of the last timed
    batch on the exact path and ABORTS above a tolerance set from its measured error (tests/test_bench_configs_gpu.py); the bf16
    error of a random latent spreads over 0.5 - 1.9e-2 of |ref|max (median 0.9e-2, profiles/engine_error_spread_r02.json), so a
    rank-dependent sample made the abort a lottery at N = 8.  Sample 1 (rank-specific) is reported beside it, not judged."""
    gen = torch.Generator().manual_seed(1234 + rank + 7919 * size)
    z = torch.randn(n_batches, B, STYLE_DIM, generator=gen)
    z[:, 0] = torch.randn(n_batches, STYLE_DIM, generator=torch.Generator().manual_seed(1234 + 7919 * size))
    return z


def measure(ctx, size, B, K, W, precision, profile_out=None, fp32_e2e=True):
    """value / e2e / roofline / clocks / parity / per-rank step times of ONE configuration"""
    sg2, dev, world, rank = ctx.sg2, ctx.dev, ctx.world, ctx.rank
    G = make_generator(sg2, size, dev, precision)
    z_host = make_latents(K + W, B, rank, size)
    z_host = z_host.pin_memory()
    z_dev = z_host.to(dev)

    def step(z):
        with torch.no_grad():
            return G([z], randomize_noise=False)[0]

    # ---- value: latents resident in HBM ---------------------------------------------------------
    sampler = ClockSampler(ctx.local).start() if rank == 0 else None
    for i in range(W):
        step(z_dev[i])
    if sampler:                          # keep the GPU under load until nvidia-smi has printed its first row
        t0 = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t0 < 3.0:
            step(z_dev[0])
            torch.cuda.synchronize()
    ctx.barrier()
    n0 = sg2._lib.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    t_begin = time.perf_counter()
    evs[0].record()
    for i in range(K):
        img = step(z_dev[W + i])
        evs[i + 1].record()
    ctx.barrier()
    t_end = time.perf_counter()
    launches = sg2._lib.launch_count() - n0
    t_dev = ctx.max_over_ranks(evs[0].elapsed_time(evs[K]) * 1e-3)
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(K))
    stats = ctx.gather([per_step[0], per_step[K // 2], per_step[-1], evs[0].elapsed_time(evs[K]) / K])
    ranks = [{"rank": r, "step_ms_min": round(float(s[0]), 3), "step_ms_median": round(float(s[1]), 3),
              "step_ms_max": round(float(s[2]), 3), "ms_per_step": round(float(s[3]), 3)} for r, s in enumerate(stats)]

    # ---- parity of the timed output: sample 0 of the last timed batch on the exact fp32 path ----
    parity = None
    if precision == "bf16" and not os.environ.get("SG2_BENCH_NO_PARITY"):      # (off only for knock-out builds, tools/knockout.sh)
        with torch.no_grad():
            G.precision = "exact"
            ref = G([z_dev[W + K - 1][:2]], randomize_noise=False)[0]
            G.precision = precision
        finite = bool(torch.isfinite(img).all().item())
        rel = float(((img[:1].float() - ref[:1]).abs().max() / ref[:1].abs().max()).item())
        rel_own = float(((img[1:2].float() - ref[1:2]).abs().max() / ref[1:2].abs().max()).item()) if B > 1 else rel
        rel_own = ctx.max_over_ranks(rel_own)
        parity = {"finite": finite, "rel_max_vs_exact_fp32": round(rel, 6), "tol": PARITY_REL_MAX,
                  "what": "sample 0 of the last timed batch (the same latent on every rank) vs the exact fp32 path (golden-pinned)",
                  "rank_specific_sample_rel_max": round(rel_own, 6)}
        ok = torch.tensor([1.0 if (finite and rel <= PARITY_REL_MAX) else 0.0], device=dev)
        if world > 1:
            ctx.dist.all_reduce(ok, op=ctx.dist.ReduceOp.MIN)
        if ok.item() < 1:
            raise RuntimeError(f"bench: the timed engine output fails parity on some rank (here: finite={finite}, rel_max={rel:.3e} "
                               f"> {PARITY_REL_MAX}) -- a number from a wrong kernel is not a result")
        del ref

    # ---- e2e: host buffers, H2D of latents and D2H of images inside the timed region -----------
    # The images leave the device the way the reference's own pipelines consume them: as uint8 (tensor2im,
    # restyle-encoder/utils/common.py:5-11), converted on the device by the repo's public `images_to_uint8`, so a
    # quarter of the bytes crosses PCIe.  The fp32 variant (every byte of the module's output) is timed too and reported
    # as e2e.fp32_images: it is bound by the host link once several ranks copy at the same time.
    copy_stream = torch.cuda.Stream(device=dev)

    def make_e2e(as_uint8):
        out_host = [torch.empty(B, 3, size, size, dtype=torch.uint8 if as_uint8 else torch.float32).pin_memory()
                    for _ in range(2)]

        def loop(n, base):
            for i in range(n):
                z = z_host[base + i].to(dev, non_blocking=True)          # H2D from pinned memory
                im = step(z)
                if as_uint8:
                    im = ctx.psp_io.images_to_uint8(im)
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
        ctx.barrier()
        t0 = time.perf_counter()
        loop(K, W)
        torch.cuda.synchronize()
        return ctx.max_over_ranks(time.perf_counter() - t0)

    t_e2e = time_e2e(True)
    t_e2e_f32 = time_e2e(False) if fp32_e2e else None

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    roofline, table = None, None
    if rank == 0 and precision == "bf16":
        table = kernel_table(ctx, G, z_dev[0], B, min(K, 10))
        roofline = roofline_of(ctx, table, size, B)
        if profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(profile_out)), exist_ok=True)
            json.dump({"batch": B, "size": size, "reps": min(K, 10), "kernels": table}, open(profile_out, "w"), indent=1)

    n_img = world * B * K
    e2e = {"value": round(n_img / t_e2e, 2), "unit": "images/s", "h2d_bytes_per_step": B * STYLE_DIM * 4,
           "d2h_bytes_per_step": B * 3 * size * size,
           "note": "pinned host latents -> Generator.forward -> images_to_uint8 (tensor2im on the device) -> uint8 images "
                   "copied back to pinned host memory (copy stream overlaps the next step)"}
    if t_e2e_f32:
        e2e["fp32_images"] = {"value": round(n_img / t_e2e_f32, 2), "d2h_bytes_per_step": B * 3 * size * size * 4,
                              "note": "same loop copying the module's fp32 output instead of uint8"}
    return {"G": G, "value": round(n_img / t_dev, 2), "ms_per_step": round(1e3 * t_dev / K, 3), "e2e": e2e, "roofline": roofline,
            "clocks": clocks, "parity": parity, "ranks": ranks, "launches": int(launches), "timed_region_s": round(t_dev, 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="synthesis", choices=["synthesis", "finetune"])
    ap.add_argument("--size", type=int, default=SIZE_DEFAULT)
    ap.add_argument("--batch", type=int, default=BATCH_DEFAULT, help="images per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip configs[3] (1024^2) and the GPU reference")
    ap.add_argument("--cpu-iters", type=int, default=5)
    ap.add_argument("--profile-out", default=None, help="write the per-kernel CUDA-event table here (json)")
    args, rest = ap.parse_known_args()
    args.warmup = max(args.warmup, 3)
    if args.workload == "finetune":
        sys.argv = [os.path.join(ROOT, "tools", "finetune_bench.py"), "--steps", str(args.steps), "--warmup", str(args.warmup)] + rest
        import runpy
        runpy.run_path(sys.argv[0], run_name="__main__")
        return 0
    if args.impl == "reference":
        return run_reference(args)

    ctx = Ctx()
    world, rank = ctx.world, ctx.rank
    B, K, W = args.batch, args.steps, args.warmup
    default_cfg = args.size == SIZE_DEFAULT and B == BATCH_DEFAULT and args.precision == "bf16"

    m = measure(ctx, args.size, B, K, W, args.precision, args.profile_out)
    G = m.pop("G")

    # the reference's GPU formulation on the same GPU (rank 0, one GPU: it is a baseline, not a scaling subject)
    gpu_ref = None
    if default_cfg and not args.no_extra and world == 1:
        try:
            gpu_ref = gpu_reference(G, args.size, B, ctx.dev)
        except Exception as e:                                   # a baseline that cannot run must not sink the bench line
            gpu_ref = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    del G
    torch.cuda.empty_cache()

    # BASELINE configs[3] in the same run: 1024^2, batch 32 per GPU
    extra = {}
    if default_cfg and not args.no_extra:
        m2 = measure(ctx, 1024, 32, K, W, "bf16", None if not args.profile_out else args.profile_out.replace(".json", "_1024.json"),
                     fp32_e2e=False)
        m2.pop("G")
        torch.cuda.empty_cache()
        extra["1024_b32"] = {"workload": workload_name(1024, 32), "value": m2["value"], "unit": "images/s",
                             "ms_per_step": m2["ms_per_step"], "e2e": m2["e2e"], "roofline": m2["roofline"],
                             "clocks": m2["clocks"], "parity": m2["parity"], "ranks": m2["ranks"],
                             "gpu_launches": m2["launches"], "global_batch": 32 * world}

    if rank != 0:
        if world > 1:
            ctx.dist.destroy_process_group()
        return 0

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ips, ts = cpu_baseline(args.size, 4, args.cpu_iters, 1)
        cpu = {"value": round(ips, 3), "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"BASELINE configs[0]: {args.size}x{args.size} forward, batch 4, fp32, 1 warm-up + {args.cpu_iters} timed "
                         f"iterations (median {1e3 * ts[len(ts) // 2]:.0f} ms/iter)"}

    act_gb = 137.5e6 * B / 1e9 if args.size == 256 else None
    line = {
        "metric": f"StyleGAN2-{args.size} images/sec", "value": m["value"], "unit": "images/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": m["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.size, B),
                   "batch_per_gpu": B, "global_batch": B * world, "precision": args.precision,
                   "l2": f"no explicit flush: each step streams >= {act_gb:.1f} GB of activations (>> 126 MB L2)" if act_gb else
                         "no explicit flush: per-step activation traffic >> 126 MB L2",
                   "parallelism": f"batch-sharded x{world}, no collective in the data path"},
        "e2e": m["e2e"], "gpu_launches": m["launches"], "clocks": m["clocks"], "parity": m["parity"], "ranks": m["ranks"],
        "timed_region_s": m["timed_region_s"],
    }
    if m["roofline"]:
        line["roofline"] = m["roofline"]
    if extra:
        line["configs"] = extra
    if gpu_ref:
        line["gpu_reference"] = gpu_ref
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

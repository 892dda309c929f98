"""CPU: the reference arm of bench.py (the oracle port timed on the host cores) prints one well-formed JSON line whose
metric / unit / workload are the ones the CUDA arm reports."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["impl"] == "reference" and d["metric"] == "StyleGAN2-256 images/sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"] == bench.workload_name(256, 64)           # the same workload as the CUDA arm names
    assert d["warmup"] >= 3 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: the driver launches the reference arm through torchrun like the CUDA arm; rank 0 alone runs and prints,
    the other ranks exit 0 without work"""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0


def test_parity_sample_is_the_same_on_every_rank():
    """bench.py judges sample 0 of the last timed batch against the exact path with a tolerance set from that sample's measured
    error: it has to be the same latent on every rank (a rank-specific one aborted the 8-GPU run), the rest of the batch must not be"""
    import torch
    import bench
    z0, z5 = bench.make_latents(6, 4, 0, 256), bench.make_latents(6, 4, 5, 256)
    assert z0.shape == (6, 4, bench.STYLE_DIM)
    assert torch.equal(z0[:, 0], z5[:, 0])
    assert not torch.equal(z0[:, 1:], z5[:, 1:])
    assert not torch.equal(bench.make_latents(6, 4, 0, 1024)[:, 0], z0[:, 0])

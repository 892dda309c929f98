"""SynthesisEngine -- host side of the whole-network bf16 tcgen05 engine (`sg2_synth_*` in the C ABI).

Replaces the synthesis loop of Generator.forward (model.py:520-533 of the reference) by ONE C call
that walks a static launch plan: style/demod tables for every layer, then per layer a TMA-fed
tcgen05/TMEM implicit-GEMM kernel with demodulation, noise, bias, leaky-relu, the next layer's
modulation and the ToRGB 1x1 projection fused into its epilogue, plus the FIR kernels of the
up-sampling layers and the RGB skip chain.  The C side allocates nothing: this class owns the
workspace (a torch uint8 tensor), hands over pointers to the fp32 master parameters and re-packs
them into the engine's bf16 layouts whenever they change.
"""
import ctypes as C
import gc
import os
import threading
import weakref

import torch

from . import _lib


class SynthesisEngine:
    def __init__(self, generator, max_batch=8, use_graph=None, training=False):
        # the module owns its engine, the engine only looks back through a weak reference: no reference cycle, so an
        # engine (plan, workspace, CUDA graphs) is torn down the moment its module dies and never by a cyclic-GC pass
        # that happens to run while some stream is capturing (destroying a graph is illegal during a global capture)
        self.G = generator
        self.lib = _lib.load()
        self.training = bool(training)       # keep every layer's output and support `backward` (frozen decoder, dL/dlatent)
        self._fwd_id = 0
        self._mod = None
        self.plan = None
        self.max_batch = 0
        self.workspace = None
        self._packed_version = None
        self._keep = []
        self.device = generator.input.input.device
        # CUDA graph per (batch, noise layout): one graph launch replaces the ~30 kernel launches of a
        # forward.  SG2_B200_GRAPH=0 switches it off; so does a caller that is not the main thread (stream capture is
        # process-global by default: an nn.DataParallel replica thread must not capture while its siblings launch).
        self.use_graph = (os.environ.get("SG2_B200_GRAPH", "1") != "0") if use_graph is None else bool(use_graph)
        self._graphs = {}
        if generator.input.input.is_cuda:
            self._ensure(max_batch)

    @property
    def G(self):
        g = self._G()
        if g is None:
            raise RuntimeError("sg2_b200 engine: its Generator no longer exists")
        return g

    @G.setter
    def G(self, generator):
        if isinstance(generator, _AdaView):
            self._view = generator           # the view has no other owner
        self._G = weakref.ref(generator)

    # -- plan -------------------------------------------------------------------------------------
    def _layer_table(self):
        G = self.G
        dev = G.input.input.device
        keep = []

        def f32(t):
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            return t.data_ptr()

        rows = []

        def styled(m, latent_index, res):
            c = m.conv
            rows.append(_lib.ConvParams(f32(c.weight), f32(c.modulation.weight), f32(c.modulation.bias),
                                        f32(m.noise.weight), f32(m.activate.bias), c.in_channel, c.out_channel,
                                        c.kernel_size, 1 if c.upsample else 0, latent_index, res))

        def rgb(m, latent_index, res):
            c = m.conv
            rows.append(_lib.ConvParams(f32(c.weight), f32(c.modulation.weight), f32(c.modulation.bias),
                                        None, f32(m.bias), c.in_channel, c.out_channel, c.kernel_size, 0,
                                        latent_index, res))

        styled(G.conv1, 0, 4)
        rgb(G.to_rgb1, 1, 4)
        i = 1
        for j in range(G.log_size - 2):
            res = 2 ** (j + 3)
            styled(G.convs[2 * j], i, res)
            styled(G.convs[2 * j + 1], i + 1, res)
            rgb(G.to_rgbs[j], i + 2, res)
            i += 2
        const = f32(G.input.input)
        taps = (G.convs[0].conv.blur.kernel.detach().float().cpu().contiguous() if len(G.convs)
                else torch.zeros(4, 4))
        self._keep = keep + [taps]
        return rows, const, taps, dev

    def _version(self):
        ps = list(self.G.parameters())
        return tuple(p._version for p in ps) + tuple(p.data_ptr() for p in ps)

    def _ensure(self, batch):
        """(re)build plan + workspace + packed weights when the batch outgrows the plan or any
        parameter changed (in-place update, .to(), load_state_dict)."""
        v = self._version()
        if self.plan is not None and batch <= self.max_batch and v == self._packed_version:
            return
        if self.plan is not None:
            self.lib.sg2_synth_destroy(self.plan)
            self.plan = None
        self._graphs = {}          # graphs hold the old plan's pointers
        rows, const, taps, dev = self._layer_table()
        if taps.shape != (4, 4):
            raise RuntimeError("sg2_b200 engine: blur kernel must have 4x4 taps (blur_kernel=[1,3,3,1])")
        arr = (_lib.ConvParams * len(rows))(*rows)
        plan = C.c_void_p()
        mb = max(batch, self.max_batch, 1)
        _lib.check(self._create_fn()(C.byref(plan), self.G.size, self.G.style_dim, mb, arr, len(rows),
                                     const, taps.numpy().ctypes.data_as(C.POINTER(C.c_float))),
                   "synth_create")
        self.plan, self.max_batch = plan, mb
        if self.training:
            _lib.check(self.lib.sg2_synth_enable_training(plan), "synth_enable_training")
            self.use_graph = False
        self._mod = None
        nbytes = self.lib.sg2_synth_workspace_bytes(plan)
        if self.workspace is None or self.workspace.numel() < nbytes or self.workspace.device != dev:
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with _lib.device_of(self.workspace):
            _lib.check(self.lib.sg2_synth_pack(self.plan, self.workspace.data_ptr(),
                                               _lib.stream_of(self.workspace)), "synth_pack")
        self._packed_version = v

    def _create_fn(self):
        return self.lib.sg2_synth_create

    def describe(self):
        buf = C.create_string_buffer(1 << 16)
        self.lib.sg2_synth_describe(self.plan, buf, len(buf))
        return buf.value.decode()

    # -- forward ----------------------------------------------------------------------------------
    def _noise_args(self, B, noise, device, into=None):
        """-> (pointer array, stride array, tensors kept alive).  `into`: static buffers to fill."""
        n_layers = self.G.num_layers
        ptrs = (C.c_void_p * n_layers)()
        strides = (C.c_int64 * n_layers)()
        keep = []
        for i in range(n_layers):
            r = 2 ** ((i + 5) // 2)
            n = noise[i] if noise is not None else None
            if into is not None:
                buf = into[i]
                if n is None:
                    buf.normal_()                                  # fresh N(0,1), model.py:283-285
                else:
                    buf.copy_(n.detach().reshape(buf.shape))
                n = buf
            else:
                if n is None:
                    n = torch.randn(B, 1, r, r, device=device, dtype=torch.float32)
                n = n.detach().float().contiguous()
                if n.data_ptr() % 16:              # the FIR kernel reads the maps through a tensor map (TMA: 16-byte aligned base)
                    n = n.clone()
            if n.numel() == B * r * r:
                strides[i] = r * r if B > 1 else 0       # one map per sample
            elif n.numel() == r * r:
                strides[i] = 0                           # one map broadcast over the batch (model.py:417-420)
            else:
                raise RuntimeError(f"noise[{i}] of shape {tuple(n.shape)} does not match [{B} or 1, 1, {r}, {r}]")
            keep.append(n)
            ptrs[i] = n.data_ptr()
        return ptrs, strides, keep

    def _noise_layout(self, B, noise):
        """per layer: number of noise maps (B or 1) the caller supplies / wants"""
        out = []
        for i in range(self.G.num_layers):
            r = 2 ** ((i + 5) // 2)
            n = noise[i] if noise is not None else None
            if n is None:
                out.append(B)
            elif n.numel() == B * r * r:
                out.append(B)
            elif n.numel() == r * r:
                out.append(1)
            else:
                raise RuntimeError(f"noise[{i}] of shape {tuple(n.shape)} does not match [{B} or 1, 1, {r}, {r}]")
        return tuple(out)

    def _call(self, lat, B, ptrs, strides, image, pooled=None, pool=0):
        """one sg2_synth_forward; with `pooled` ([B,3,size/pool,size/pool] fp32) the last launch also writes the face-pooled
        image (sg2_synth_set_pooled_output) and `image` may be None (pooled image only)"""
        with _lib.device_of(lat):
            if pool:
                _lib.check(self.lib.sg2_synth_set_pooled_output(self.plan, pooled.data_ptr(), pool, 0 if image is None else 1),
                           "synth_set_pooled_output")
            try:
                _lib.check(self.lib.sg2_synth_forward(self.plan, self.workspace.data_ptr(), lat.data_ptr(), B, ptrs, strides,
                                                      None if image is None else image.data_ptr(), _lib.stream_of(lat)),
                           "synth_forward")
            finally:
                if pool:
                    self.lib.sg2_synth_set_pooled_output(self.plan, None, 0, 1)

    def _pool_request(self, B, dev):
        """face_pool folded into the forward (psp_io.decode_pooled sets Generator._pool_request = factor): -> factor or 0"""
        f = getattr(self._G(), "_pool_request", None) if not self.training else None
        if f in (2, 4) and self.G.size % (4 * f) == 0:
            return f
        return 0

    @torch.no_grad()
    def synthesize(self, latent, noise, graph=None, z=None, want_latent=True):
        """latent [B, n_latent, style_dim]; noise: list (num_layers) of [B or 1, 1, r, r] tensors or
        None entries (fresh N(0,1) is drawn, model.py:283-285).  Returns the image [B,3,size,size].
        z [B, style_dim] instead of latent (latent=None): the mapping network and the broadcast over the layers
        (model.py:482-483,503-506) run INSIDE the CUDA graph; returns (image, latent)."""
        G = self.G
        from_z = z is not None
        if from_z:
            _lib.require_cuda(z, "z")
            if latent is None:
                if not (self.use_graph if graph is None else graph) or torch.cuda.is_current_stream_capturing() or \
                        threading.current_thread() is not threading.main_thread() or z.shape[0] == 0:
                    w = G.style(z)
                    latent = w.unsqueeze(1).repeat(1, G.n_latent, 1)
                    return self.synthesize(latent, noise, graph=False), latent
                latent = z                       # batch size / device below
        _lib.require_cuda(latent, "latent")
        B = latent.shape[0]
        self._ensure(B)
        out_dtype = G.input.input.dtype
        dev = latent.device
        use_graph = self.use_graph if graph is None else graph
        if use_graph and threading.current_thread() is not threading.main_thread():
            use_graph = False
        if dev != self.device:
            raise RuntimeError(f"sg2_b200 engine: planned for {self.device}, called with latents on {dev} "
                               "(Generator.engine() re-plans after the module moves)")
        if B == 0:
            return torch.empty(0, 3, G.size, G.size, device=dev, dtype=out_dtype)
        pool = self._pool_request(B, dev)
        if not use_graph or torch.cuda.is_current_stream_capturing():
            lat = latent.detach().float().contiguous()
            ptrs, strides, keep = self._noise_args(B, noise, dev)
            if pool:                             # pooled image only: the full-resolution one is never written
                image = torch.empty(B, 3, G.size // pool, G.size // pool, device=dev, dtype=torch.float32)
                self._call(lat, B, ptrs, strides, None, pooled=image, pool=pool)
            else:
                image = torch.empty(B, 3, G.size, G.size, device=dev, dtype=torch.float32)
                self._call(lat, B, ptrs, strides, image)
            return image if out_dtype == torch.float32 else image.to(out_dtype)

        layout = self._noise_layout(B, noise)
        key = (B, layout, dev.index, from_z, pool)
        entry = self._graphs.get(key)
        if entry is None:
            s_lat = torch.empty(B, G.n_latent, G.style_dim, device=dev, dtype=torch.float32)
            s_z = torch.empty(B, G.style_dim, device=dev, dtype=z.dtype) if from_z else None

            def head():                      # mapping network + broadcast, part of the graph when the caller passes z
                if from_z:
                    s_lat.copy_(G.style(s_z).unsqueeze(1).expand(-1, G.n_latent, -1))
            # all static noise maps live in ONE flat buffer: a fully randomised forward refills it with one launch
            shapes = [(nb, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i, nb in enumerate(layout)]
            sizes = [s[0] * s[2] * s[3] for s in shapes]
            s_flat = torch.empty(sum(sizes), device=dev, dtype=torch.float32)
            s_noise, o = [], 0
            for shp, n in zip(shapes, sizes):
                s_noise.append(s_flat[o:o + n].view(shp))
                o += n
            s_img = torch.empty(B, 3, G.size // (pool or 1), G.size // (pool or 1), device=dev, dtype=torch.float32)
            (s_z if from_z else s_lat).copy_(z if from_z else latent)
            ptrs, strides, _ = self._noise_args(B, noise, dev, into=s_noise)
            n0 = _lib.launch_count()
            head()

            def run():
                if pool:
                    self._call(s_lat, B, ptrs, strides, None, pooled=s_img, pool=pool)
                else:
                    self._call(s_lat, B, ptrs, strides, s_img)
            run()                                                   # warm-up outside capture (lazy attributes, descriptors)
            n_launch = _lib.launch_count() - n0
            torch.cuda.current_stream(dev).synchronize()
            g = torch.cuda.CUDAGraph()
            # Nothing may free device objects while the stream captures: collect garbage now, keep the collector off for
            # the few milliseconds of the capture (a cyclic-GC pass that finalises somebody's old CUDA graph inside it
            # invalidates the capture), and let other threads go on using CUDA (thread-local error mode).
            gc.collect()
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    head()
                    run()
            finally:
                if gc_was_on:
                    gc.enable()
            entry = {"graph": g, "lat": s_lat, "z": s_z, "noise": s_noise, "flat": s_flat, "img": s_img, "launches": n_launch,
                     "args": (ptrs, strides), "src": [None] * len(s_noise)}
            self._graphs[key] = entry
        if from_z:
            entry["z"].copy_(z)
        else:
            entry["lat"].copy_(latent)
        self._stage_noise(entry, noise)
        entry["graph"].replay()
        self.lib.sg2_note_launches(entry["launches"])              # kernels launched by the graph replay
        s_img = entry["img"]
        image = s_img.clone() if out_dtype == torch.float32 else s_img.to(out_dtype)
        if from_z:                                                  # a copy: the static buffer is overwritten by the next call
            return image, (entry["lat"].to(z.dtype, copy=True) if want_latent else None)
        return image

    @staticmethod
    def _stage_noise(entry, noise):
        """Bring the graph's static noise maps up to date: one normal_() over the flat buffer when every
        layer draws fresh noise (model.py:283-285), no copy at all for caller tensors that have not
        changed since they were last staged (the registered `noises` buffers of randomize_noise=False)."""
        bufs, src = entry["noise"], entry["src"]
        if noise is None or all(n is None for n in noise):
            entry["flat"].normal_()
            for i in range(len(src)):
                src[i] = None
            return
        for i, buf in enumerate(bufs):
            n = noise[i]
            if n is None:
                buf.normal_()
                src[i] = None
                continue
            # same tensor OBJECT (weak reference: a dead source can never match, so a recycled address is
            # not mistaken for it) at the same version -> the staged copy is still current
            prev = src[i]
            if prev is None or prev[0]() is not n or prev[1] != n._version:
                buf.copy_(n.detach().reshape(buf.shape))
                src[i] = (weakref.ref(n), n._version)

    # -- training: forward that keeps the activations + backward w.r.t. the latents ---------------
    def _modulation_table(self):
        """per plan row: (latent index, cin, W * scale [cin, style_dim]) -- the EqualLinear of every modulation
        (model.py:152-155, 222), used to turn dL/dstyle into dL/dlatent"""
        if self._mod is None:
            G = self.G
            rows = []

            def add(conv, idx):
                m = conv.modulation
                rows.append((idx, conv.in_channel, (m.weight.detach().float() * m.scale).contiguous()))
            add(G.conv1.conv, 0)
            add(G.to_rgb1.conv, 1)
            i = 1
            for j in range(G.log_size - 2):
                add(G.convs[2 * j].conv, i)
                add(G.convs[2 * j + 1].conv, i + 1)
                add(G.to_rgbs[j].conv, i + 2)
                i += 2
            self._mod = rows
        return self._mod

    def _train_graphs(self, latent, noise):
        """CUDA graphs of the training forward and of the backward (+ the dL/dstyle -> dL/dlatent products) for this batch size
        and noise layout, or None when capture is not possible here.  At B = 8 (the ReStyle batch) a decoder pass is ~30 + ~60
        short launches: replayed as two graphs they run back to back (forward 1.40 -> 1.15 ms)."""
        if os.environ.get("SG2_B200_TRAIN_GRAPH", "1") == "0" or torch.cuda.is_current_stream_capturing():
            return None
        G = self.G
        B, dev = latent.shape[0], latent.device
        layout = self._noise_layout(B, noise)
        key = ("train", B, layout, dev.index)
        e = self._graphs.get(key)
        if e is not None:
            return e
        if threading.current_thread() is not threading.main_thread():
            return None                          # capture only from the main thread (replay is fine anywhere)
        mod = self._modulation_table()
        total = B * sum(cin for _, cin, _ in mod)
        s_lat = torch.empty(B, G.n_latent, G.style_dim, device=dev, dtype=torch.float32)
        shapes = [(nb, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2)) for i, nb in enumerate(layout)]
        sizes = [sh[0] * sh[2] * sh[3] for sh in shapes]
        s_flat = torch.empty(sum(sizes), device=dev, dtype=torch.float32)
        s_noise, o = [], 0
        for shp, n in zip(shapes, sizes):
            s_noise.append(s_flat[o:o + n].view(shp))
            o += n
        s_img = torch.empty(B, 3, G.size, G.size, device=dev, dtype=torch.float32)
        s_gimg = torch.zeros(B, 3, G.size, G.size, device=dev, dtype=torch.float32)
        s_gs = torch.empty(total, device=dev, dtype=torch.float32)
        s_acc = torch.empty(G.n_latent, B, G.style_dim, device=dev, dtype=torch.float32)
        s_out = torch.empty(B, G.n_latent, G.style_dim, device=dev, dtype=torch.float32)
        s_lat.copy_(latent.detach())
        ptrs, strides, _ = self._noise_args(B, noise, dev, into=s_noise)

        def fwd():
            self._call(s_lat, B, ptrs, strides, s_img)

        def bwd():
            with _lib.device_of(s_gimg):
                _lib.check(self.lib.sg2_synth_backward(self.plan, self.workspace.data_ptr(), B, ptrs, strides, s_gimg.data_ptr(),
                                                       s_gs.data_ptr(), _lib.stream_of(s_gimg)), "synth_backward")
            s_acc.zero_()
            off = 0
            for idx, cin, w in mod:
                s_acc[idx].addmm_(s_gs[off:off + B * cin].view(B, cin), w)
                off += B * cin
            s_out.copy_(s_acc.permute(1, 0, 2))

        n0 = _lib.launch_count()
        fwd()                                    # warm-up outside capture (descriptors, lazy attributes, library handles)
        n1 = _lib.launch_count()
        bwd()
        n2 = _lib.launch_count()
        torch.cuda.current_stream(dev).synchronize()
        gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(gf, capture_error_mode="thread_local"):
                fwd()
            with torch.cuda.graph(gb, capture_error_mode="thread_local"):
                bwd()
        finally:
            if gc_was_on:
                gc.enable()
        e = {"fwd": gf, "bwd": gb, "lat": s_lat, "noise": s_noise, "flat": s_flat, "img": s_img, "gimg": s_gimg, "gs": s_gs,
             "acc": s_acc, "out": s_out, "args": (ptrs, strides), "src": [None] * len(s_noise),
             "launches_fwd": n1 - n0, "launches_bwd": n2 - n1}
        self._graphs[key] = e
        return e

    def train_forward(self, latent, noise):
        """-> (image fp32 [B,3,size,size], state for train_backward)"""
        assert self.training
        G = self.G
        _lib.require_cuda(latent, "latent")
        B = latent.shape[0]
        self._ensure(B)
        e = self._train_graphs(latent, noise) if B > 0 else None
        if e is not None:
            e["lat"].copy_(latent.detach())
            self._stage_noise(e, noise)
            e["fwd"].replay()
            self.lib.sg2_note_launches(e["launches_fwd"])
            self._fwd_id += 1
            return e["img"].clone(), (self._fwd_id, B, e, None, None)
        lat = latent.detach().float().contiguous()
        ptrs, strides, keep = self._noise_args(B, noise, latent.device)
        image = torch.empty(B, 3, G.size, G.size, device=latent.device, dtype=torch.float32)
        self._call(lat, B, ptrs, strides, image)
        self._fwd_id += 1
        return image, (self._fwd_id, B, ptrs, strides, keep)

    def train_backward(self, state, grad_image):
        """dL/dlatent [B, n_latent, style_dim] fp32 from dL/dimage; the forward that produced `state` must be the last
        one this engine ran (the kept activations are overwritten by the next forward)"""
        fwd_id, B, ptrs, strides, keep = state
        if fwd_id != self._fwd_id:
            raise RuntimeError("sg2_b200 training engine: another forward pass ran on this Generator before backward(); "
                               "the kept activations are gone.  Call backward() after each forward (as the ReStyle "
                               "coaches do), or set SG2_B200_TRAIN_ENGINE=0 to use the autograd path")
        if isinstance(ptrs, dict):               # graph entry of _train_graphs
            e = ptrs
            e["gimg"].copy_(grad_image.detach())
            e["bwd"].replay()
            self.lib.sg2_note_launches(e["launches_bwd"])
            return e["out"].clone()
        G = self.G
        g = grad_image.detach().float().contiguous()
        mod = self._modulation_table()
        total = B * sum(cin for _, cin, _ in mod)
        gs = torch.empty(total, device=g.device, dtype=torch.float32)
        with _lib.device_of(g):
            _lib.check(self.lib.sg2_synth_backward(self.plan, self.workspace.data_ptr(), B, ptrs, strides, g.data_ptr(),
                                                   gs.data_ptr(), _lib.stream_of(g)), "synth_backward")
        acc = torch.zeros(G.n_latent, B, G.style_dim, device=g.device, dtype=torch.float32)
        off = 0
        for idx, cin, w in mod:
            acc[idx].addmm_(gs[off:off + B * cin].view(B, cin), w)
            off += B * cin
        return acc.permute(1, 0, 2).contiguous()

    def __reduce__(self):
        raise TypeError("SynthesisEngine holds device handles and cannot be pickled; copy the Generator instead "
                        "(its copies re-plan lazily)")

    def __del__(self):
        try:
            if self.plan is not None:
                self.lib.sg2_synth_destroy(self.plan)
        except Exception:
            pass


class SynthesisFunction(torch.autograd.Function):
    """The frozen decoder as ONE autograd node: forward = the engine's launch plan (activations kept), backward = the
    plan walked in reverse (csrc/synth_train.cu).  Gradient flows to `latent` only."""

    @staticmethod
    def forward(ctx, latent, engine, noise):
        image, state = engine.train_forward(latent, noise)
        ctx.engine, ctx.state = engine, state
        ctx.lat_dtype = latent.dtype
        return image

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_image):
        g = ctx.engine.train_backward(ctx.state, grad_image)
        return g.to(ctx.lat_dtype), None, None


class _AdaView:
    """what SynthesisEngine reads from a Generator, answered by a stylegan2_ada SynthesisNetwork"""

    def __init__(self, syn):
        import types
        self._syn = weakref.ref(syn)          # the network owns the engine, the engine owns this view: no cycle
        self.size, self.style_dim = syn.img_resolution, syn.w_dim
        self.log_size = syn.img_resolution_log2
        self.n_latent = 2 * self.log_size - 2            # rows of ws the layers read (the reference allocates two more)
        self.num_layers = 2 * (self.log_size - 2) + 1
        self.input = types.SimpleNamespace(input=syn.first_block.const)

    @property
    def syn(self):
        s = self._syn()
        if s is None:
            raise RuntimeError("sg2_b200 engine: its SynthesisNetwork no longer exists")
        return s

    def parameters(self):
        return self.syn.parameters()


class AdaSynthesisEngine(SynthesisEngine):
    """The same launch plan for the stylegan2_ada decoder (`sg2_synth_create_ada`): conv -> SmoothUpsample ordering,
    clamps, no equalised-lr conv scale.  Inference only (the differentiable composition stays in stylegan2_ada/)."""

    def __init__(self, synthesis, max_batch=8, use_graph=None):
        self._view = _AdaView(synthesis)
        super().__init__(self._view, max_batch=max_batch, use_graph=use_graph)

    def _create_fn(self):
        return self.lib.sg2_synth_create_ada

    def _layer_table(self):
        syn = self.G.syn
        dev = syn.first_block.const.device
        keep = []

        def f32(t):
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            return t.data_ptr()

        rows = []
        PAD = 32      # narrowest channel count of the tensor-core kernels (one 64-byte swizzle row)

        # Layers narrower than 32 channels (the 1024^2 block of the ADA decoder has 16) run zero-padded to 32: padded input
        # channels get a zero style (affine weight and bias rows = 0) and zero weights, padded output channels zero weights, a
        # zero bias and -- through the consumer's zero style -- are stored as exact zeros, so the arithmetic of the real
        # channels is untouched; it costs 2x the bytes on those layers only.
        def padded(t, shape):
            t = t.detach().float()
            if tuple(t.shape) == tuple(shape):
                return t
            out = torch.zeros(shape, device=t.device, dtype=torch.float32)
            out[tuple(slice(0, n) for n in t.shape)] = t
            return out

        def styled(m, latent_index, res, up):
            cout, cin = m.weight.shape[:2]
            co, ci = max(cout, PAD), max(cin, PAD)
            rows.append(_lib.ConvParams(f32(padded(m.weight, (co, ci, 3, 3))), f32(padded(m.affine.weight, (ci, m.affine.weight.shape[1]))),
                                        f32(padded(m.affine.bias, (ci,))), f32(m.noise_strength), f32(padded(m.bias, (co,))),
                                        ci, co, 3, 1 if up else 0, latent_index, res))

        def rgb(m, latent_index, res):
            cout, cin = m.weight.shape[:2]
            ci = max(cin, PAD)
            # the ToRGB weight gain 1 / sqrt(cin) is applied by the engine from the PADDED width: pre-compensate
            w = padded(m.weight, (cout, ci, 1, 1)) * float((ci / cin) ** 0.5)
            rows.append(_lib.ConvParams(f32(w), f32(padded(m.affine.weight, (ci, m.affine.weight.shape[1]))),
                                        f32(padded(m.affine.bias, (ci,))), None, f32(m.bias), ci, cout, 1, 0, latent_index, res))

        fb = syn.first_block
        styled(fb.conv1, 0, 4, False)
        rgb(fb.torgb, 1, 4)
        for n, blk in enumerate(syn.blocks):
            res = 8 << n
            styled(blk.conv0, 2 * n + 1, res, True)
            styled(blk.conv1, 2 * n + 2, res, False)
            rgb(blk.torgb, 2 * n + 3, res)
        const = f32(fb.const)
        taps = (syn.blocks[0].resampler.kernel.detach().float().cpu().reshape(4, 4).contiguous() if len(syn.blocks)
                else torch.zeros(4, 4))
        self._keep = keep + [taps]
        return rows, const, taps, dev

import importlib, sys, torch
sys.path.insert(0, "/root/repo")
sg2 = importlib.import_module("stylegan-for-facerec_b200")
taps = sg2.make_kernel([1, 3, 3, 1]).cuda()
for (B, c, r) in ((4096, 256, 8), (4096, 512, 8), (4096, 32, 16)):
    x = torch.randn(B, c, r, r, device="cuda")
    for name, fn in (("up2", lambda: sg2.upfirdn2d(x, taps * 4, up=2, pad=(2, 1))), ("down2", lambda: sg2.upfirdn2d(x, taps, down=2, pad=(1, 1))),
                     ("lrelu", lambda: sg2.fused_leaky_relu(x, torch.randn(c, device="cuda")))):
        try:
            y = fn(); torch.cuda.synchronize(); print(B, c, r, name, "ok", tuple(y.shape), flush=True)
        except Exception as e:
            print(B, c, r, name, "FAILED", repr(e)[:300], flush=True); raise
    xb = torch.randn(B, c, r + 1, r + 1, device="cuda")
    y = sg2.upfirdn2d(xb, taps * 4, pad=(1, 1)); torch.cuda.synchronize(); print(B, c, r, "blur ok", flush=True)

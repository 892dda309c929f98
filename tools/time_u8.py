import importlib, sys, torch, time
sys.path.insert(0, "/root/repo")
io = importlib.import_module("stylegan-for-facerec_b200.psp_io")
x = torch.randn(32, 3, 1024, 1024, device="cuda")
for _ in range(3): y = io.images_to_uint8(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): y = io.images_to_uint8(x)
b.record(); torch.cuda.synchronize()
print("images_to_uint8 32x3x1024x1024: %.3f ms" % (a.elapsed_time(b) / 10))
ref = ((x.cpu().numpy() + 1) / 2).clip(0, 1) * 255
print("match", (y.cpu().numpy() == ref.astype("uint8")).mean())

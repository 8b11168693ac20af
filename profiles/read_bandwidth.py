import torch, time
n = 1953736209
v = torch.rand(n, dtype=torch.float64, device="cuda")
for f, name in ((lambda: v.sum(), "torch sum (read 15.6 GB)"), (lambda: v.max(), "torch max")):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, f"{ms:.3f} ms  {8 * n / ms / 1e6:.0f} GB/s")
w = torch.empty_like(v)
w.copy_(v); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): w.copy_(v)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("torch copy (read + write)", f"{ms:.3f} ms  {16 * n / ms / 1e6:.0f} GB/s")

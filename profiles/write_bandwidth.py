"""Write-only / read-only / copy bandwidth of this GPU with torch kernels (the K_e round trip of the two-pass assembly
writes 36.9 GB and reads them back): python profiles/write_bandwidth.py"""
import torch
n = 4 * 1024 ** 3 // 8 * 8          # 32 GiB of doubles
a = torch.empty(n, dtype=torch.float64, device="cuda")
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t = timed(lambda: a.fill_(1.0)); print(f"fill (write only) {a.numel() * 8 / t / 1e6:8.0f} GB/s")
t = timed(lambda: a.zero_()); print(f"memset            {a.numel() * 8 / t / 1e6:8.0f} GB/s")
t = timed(lambda: a.sum()); print(f"sum (read only)   {a.numel() * 8 / t / 1e6:8.0f} GB/s")
b = torch.empty(n // 2, dtype=torch.float64, device="cuda")
t = timed(lambda: b.copy_(a[: n // 2])); print(f"copy (read+write) {b.numel() * 16 / t / 1e6:8.0f} GB/s")

"""Write-only / read-only / copy bandwidth of the B200's HBM with plain torch kernels (context for the roofline of a
92 %-writes kernel: the denominator in bench.py is the driver's read+write COPY figure)."""
import torch
n = 1 << 30  # 8 GiB of float64
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.0)); print(f"fill (write only)  {8*n/ms/1e6:8.1f} GB/s")
ms = t(lambda: y.copy_(x)); print(f"copy (read+write)  {16*n/ms/1e6:8.1f} GB/s")
ms = t(lambda: x.sum()); print(f"sum  (read only)   {8*n/ms/1e6:8.1f} GB/s")
ms = t(lambda: torch.add(x, 1.0, out=x)); print(f"x += 1 (r+w same)  {16*n/ms/1e6:8.1f} GB/s")

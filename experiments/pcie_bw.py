import torch, time
n = 1 << 28  # 2 GiB of float64
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    print(name, round(8 * n / best / 1e9, 1), "GB/s")
# chunked D2H as the host path does it: 184 MB pieces on one stream
k = (184 << 20) // 8
t = time.perf_counter()
for i in range(0, n - k, k):
    h[i:i + k].copy_(d[i:i + k], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print("D2H in 184 MB pieces", round(8 * (n // k) * k / dt / 1e9, 1), "GB/s")

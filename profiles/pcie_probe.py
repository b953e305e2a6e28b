import torch, time
n = 1_800_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def chunks():
    with torch.cuda.stream(s1):
        for k in range(0, n, 47_000_000): d[k:k+47_000_000].copy_(h[k:k+47_000_000], non_blocking=True)
a = t(h2d); print("H2D alone GB/s", n / a / 1e9)
b = t(both); print("H2D with concurrent D2H(0.9GB): H2D-equivalent GB/s", n / b / 1e9)
c = t(chunks); print("H2D in 47MB chunks GB/s", n / c / 1e9)
import os; print("cpus", os.cpu_count())

"""Write / read / copy bandwidth microbenchmark (context for the Jacobian kernel's roofline; not a test)."""
import torch
n = 444 * 1024 * 1024 // 8
a = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
def t(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
nb = n * 8
print("fill   (write only) GB/s", nb / t(lambda: a.fill_(1.0)) / 1e9)
print("sum    (read only)  GB/s", nb / t(lambda: a.sum()) / 1e9)
print("copy   (r+w bytes)  GB/s", 2 * nb / t(lambda: b.copy_(a)) / 1e9)
c = torch.empty(n // 5, dtype=torch.float64, device="cuda")
print("copy 1/5 read + full write-ish: mixed 20/80 not directly available")

"""Pure-write / pure-read / copy HBM bandwidth on this box (context for the write-heavy kernels' rooflines)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
def t(f, reps=10):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(f"fill (write 2 GiB): {2 * n / t(lambda: a.zero_()) / 1e6:.0f} GB/s")
print(f"sum  (read 2 GiB) : {2 * n / t(lambda: a.view(torch.int16).max()) / 1e6:.0f} GB/s")
print(f"copy (r+w 4 GiB)  : {4 * n / t(lambda: b.copy_(a)) / 1e6:.0f} GB/s")

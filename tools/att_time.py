"""Time sunb_attention at the bench shapes (2500 images) for the library named by SUNB200_LIB."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "few-shot-vit_b200"))
from sunb200 import native as N
lib, st = N.lib(), N.current_stream()
for (S, d, dp) in ((100, 42, 48), (25, 85, 96)):
    B, heads = 2500, 6
    qkv = (torch.randn(B * S, 3 * heads * dp, device="cuda") * 0.5).bfloat16()
    out = torch.empty(B * S, heads * dp, device="cuda", dtype=torch.bfloat16)
    f = lambda: lib.sunb_attention(qkv.data_ptr(), out.data_ptr(), B, S, d, dp, heads, 3 * heads * dp, heads * dp, st)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(f"S={S} d={d}: {e0.elapsed_time(e1) * 100:.1f} us", end="; ")
print(os.environ.get("SUNB200_LIB", "default")[-8:])

"""Time the attention cores at the bench chunk size (B = 2500): python tools/att_time.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-vit_b200"))
import torch
from sunb200 import native as N
lib, st = N.lib(), N.current_stream()
B = 2500
for S, d, dp in ((100, 42, 48), (25, 85, 96)):
    qkv = torch.randn(B * S, 18 * dp, device="cuda").bfloat16()
    out = torch.empty(B * S, 6 * dp, device="cuda", dtype=torch.bfloat16)
    f = lambda: N.check(lib.sunb_attention(qkv.data_ptr(), out.data_ptr(), B, S, d, dp, 6, 18 * dp, 6 * dp, st), "att")
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nbytes = B * S * 24 * dp * 2
    print(f"S={S} d={d}: {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s")

"""Time the attention backward: the product's tcgen05 kernel (padded heads) and the warp-MMA cross-check kernel of the test
library (the reference's packed heads) at a batch size: python tools/attb_time.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-vit_b200"))
import torch
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sunb200 import native as N
from gpu_helpers import check_lib, check_call
lib, st = N.lib(), N.current_stream()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 480
for S, d, ds in ((100, 42, 42), (100, 42, 48), (25, 85, 85), (25, 85, 96)):
    ld3, ldi = (18 * ds + 15) // 16 * 16, (6 * ds + 15) // 16 * 16
    qkv = torch.randn(B * S, ld3, device="cuda").bfloat16()
    dout = torch.randn(B * S, ldi, device="cuda").bfloat16()
    dqkv = torch.empty(B * S, ld3, device="cuda", dtype=torch.bfloat16)
    if ds == d:
        f = lambda: check_call(check_lib().sunb_check_attention_backward(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, S, d, 6, ld3, ldi, st), "check attb")
    else:
        f = lambda: N.check(lib.sunb_attention_backward(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, S, d, ds, 6, ld3, ldi, st), "attb")
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    flops = 5 * 2.0 * B * 6 * S * S * d
    print(f"B={B} S={S} d={d} ds={ds}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s (useful)")

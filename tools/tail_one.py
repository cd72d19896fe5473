"""The fused stage-1 block tail alone (for ncu): python tools/tail_one.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-vit_b200"))
import torch
from sunb200 import native as N, packing
lib, st = N.lib(), N.current_stream()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
x = torch.randn(B * 400, 128, device="cuda").bfloat16()
xo = torch.empty_like(x)
h1 = torch.randn(B * 400, 256, device="cuda").bfloat16()
blob = packing._convmlp_tail_blob(torch.randn(256, 32, 3, 3) * 0.06, torch.randn(128, 256) * 0.06).bfloat16().cuda().contiguous()
for _ in range(3):
    N.check(lib.sunb_convmlp_tail(h1.data_ptr(), blob.data_ptr(), x.data_ptr(), xo.data_ptr(), B, 0, st), "tail")
torch.cuda.synchronize()

"""One shape of the attention core for ncu: python tools/att_one.py S d dp"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-vit_b200"))
import torch
from sunb200 import native as N
S, d, dp = (int(a) for a in sys.argv[1:4])
lib, st = N.lib(), N.current_stream()
B = 2500
qkv = torch.randn(B * S, 18 * dp, device="cuda").bfloat16()
out = torch.empty(B * S, 6 * dp, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    N.check(lib.sunb_attention(qkv.data_ptr(), out.data_ptr(), B, S, d, dp, 6, 18 * dp, 6 * dp, st), "att")
torch.cuda.synchronize()

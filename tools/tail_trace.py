"""Event timeline of the fused stage-1 block tail (CTA 0, third work item) from a -DSUNB_TAIL_TRACE build:
tools/build_variants.sh trace && SUNB200_LIB=tools/_ab/libsunb200_trace.so python tools/tail_trace.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-vit_b200"))
import torch
from sunb200 import native as N, packing
lib, st = N.lib(), N.current_stream()
B = 2500
x = torch.randn(B * 400, 128, device="cuda").bfloat16(); xo = torch.empty_like(x)
h1 = torch.randn(B * 400, 256, device="cuda").bfloat16()
w2 = torch.randn(256, 32, 3, 3, device="cuda") * 0.06; w3 = torch.randn(128, 256, device="cuda") * 0.06
blob = packing._convmlp_tail_blob(w2, w3).bfloat16().contiguous()
for _ in range(3):
    N.check(lib.sunb_convmlp_tail(h1.data_ptr(), blob.data_ptr(), x.data_ptr(), xo.data_ptr(), B, 0, st), "tail")
torch.cuda.synchronize()
raw = C.CDLL(N.LIB_PATH)
buf = (C.c_longlong * (16 * 12))()
assert raw.sunb_tail_trace_read(buf) == 0
t = [[buf[g * 12 + e] for e in range(12)] for g in range(16)]
t0 = t[0][0]
print("group | issuer1: start atom_ready w2+d2_ready issued | issuer2: start ready issued | GELU(t=0): wait_start d2_full h2_signalled")
for g in range(8):
    r = [v - t0 for v in t[g]]
    print(f"  {g}   {r[0]:6d} {r[1]:6d} {r[2]:6d} {r[3]:6d}   | {r[4]:6d} {r[5]:6d} {r[6]:6d}   | {r[7]:6d} {r[8]:6d} {r[9]:6d}")
print("epilogue warp 20: wait_start d3_full stores_done:", [v - t0 for v in t[8][:3]])

"""Stage-1 block at the bench chunk size: conv1 GEMM + fused tail vs the three stand-alone kernels.  python tools/convmlp_time.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-vit_b200"))
import torch
from sunb200 import native as N, packing
lib, st = N.lib(), N.current_stream()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
dev = "cuda"
x = torch.randn(B * 400, 128, device=dev).bfloat16()
xo = torch.empty_like(x)
h1 = torch.randn(B * 400, 256, device=dev).bfloat16()
h2 = torch.empty_like(h1)
w1 = (torch.randn(256, 128, device=dev) * 0.09).bfloat16()
b1 = torch.randn(256, device=dev) * 0.1
w2 = torch.randn(256, 32, 3, 3, device=dev) * 0.06
w3 = torch.randn(128, 256, device=dev) * 0.06
wg = packing._grouped_taps(w2).bfloat16().contiguous()
w3b = w3.bfloat16().contiguous()
blob = packing._convmlp_tail_blob(w2, w3).bfloat16().contiguous()

def desc(M, Nn, K, A, lda, W, ldw, o, ldc, bias=None, act=0, resid=None):
    d = N.GemmDesc()
    d.M, d.N, d.K, d.taps, d.groups = M, Nn, K, 1, 1
    d.A, d.lda, d.Wt, d.ldw = A.data_ptr(), lda, W.data_ptr(), ldw
    d.bias, d.bias_mod, d.act = N.ptr(bias), 1, act
    if resid is not None:
        d.resid, d.ldr = resid.data_ptr(), resid.shape[-1]
    d.rows_per_img = 1
    d.out, d.ldc = o.data_ptr(), ldc
    return d
d1 = desc(B * 400, 256, 128, x, 128, w1, 128, h1, 256, bias=b1, act=2)
d3 = desc(B * 400, 128, 256, h2, 256, w3b, 256, xo, 128, resid=x)

def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
conv1 = lambda: N.check(lib.sunb_gemm(C.byref(d1), 0, st), "conv1")
gconv = lambda: N.check(lib.sunb_gconv3x3(h1.data_ptr(), 256, wg.data_ptr(), h2.data_ptr(), 256, None, 0, None, 0, B, 2, 0, st), "gconv")
conv3 = lambda: N.check(lib.sunb_gemm(C.byref(d3), 0, st), "conv3")
tail = lambda: N.check(lib.sunb_convmlp_tail(h1.data_ptr(), blob.data_ptr(), x.data_ptr(), xo.data_ptr(), B, 0, st), "tail")
print(f"B={B}: conv1 {t(conv1):.1f} us, gconv {t(gconv):.1f} us, conv3 {t(conv3):.1f} us, fused tail {t(tail):.1f} us")
blk = lambda: (conv1(), tail())
tb = t(blk)
print(f"B={B}: conv1 + tail back to back {tb:.1f} us = {tb * 2500 / B:.1f} us per 2500 images")

"""Residual-add GEMMs of the attention blocks at 2500 images (proj and conv3 shapes): python tools/gemm_resid_time.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-vit_b200"))
import torch
from sunb200 import native as N
lib, st = N.lib(), N.current_stream()
def run(name, M, Nn, K, resid=True):
    A = torch.randn(M, K, device="cuda").bfloat16(); W = (torch.randn(Nn, K, device="cuda") * 0.05).bfloat16()
    x = torch.randn(M, Nn, device="cuda").bfloat16(); out = torch.empty_like(x)
    d = N.GemmDesc()
    d.M, d.N, d.K, d.taps, d.groups = M, Nn, K, 1, 1
    d.A, d.lda, d.Wt, d.ldw = A.data_ptr(), K, W.data_ptr(), K
    d.bias_mod, d.rows_per_img = 1, 1
    d.out, d.ldc = out.data_ptr(), Nn
    if resid:
        d.resid, d.ldr = x.data_ptr(), Nn
    f = lambda: N.check(lib.sunb_gemm(C.byref(d), 0, st), name)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ref = A[:512].float() @ W.float().t() + (x[:512].float() if resid else 0)
    err = float((out[:512].float() - ref).norm() / ref.norm())
    print(f"{name}: {ms*1e3:.1f} us  {2.0*M*Nn*K/ms/1e9:.0f} TFLOP/s  rel err {err:.1e}")
run("stage2 proj  M=250000 N=256 K=288", 250000, 256, 288)
run("stage3 proj  M=62500  N=512 K=576", 62500, 512, 576)
run("stage3 conv3 M=62500  N=512 K=2048", 62500, 512, 2048)
run("stage2 qkv   M=250000 N=864 K=256 (no residual)", 250000, 864, 256, resid=False)
run("stage3 qkv   M=62500  N=1728 K=512 (no residual)", 62500, 1728, 512, resid=False)

"""Stage-2 MLP at the bench chunk size: fused kernel vs the two GEMMs it replaces (+ a small correctness check).
python tools/mlp_time.py [B]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "few-shot-vit_b200"))
import torch
from sunb200 import native as N
lib, st = N.lib(), N.current_stream()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
M, Cc, H = B * 100, 256, 1024
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(3)
x = torch.randn(M, Cc, device=dev, generator=g).bfloat16()
w1 = (torch.randn(H, Cc, device=dev, generator=g) * 0.06).bfloat16()
b1 = torch.randn(H, device=dev, generator=g) * 0.1
w3 = (torch.randn(Cc, H, device=dev, generator=g) * 0.03).bfloat16()
hid = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
out_a, out_b = torch.empty_like(x), torch.empty_like(x)

def desc(Mm, Nn, K, A, W, o, bias=None, act=0, resid=None):
    d = N.GemmDesc()
    d.M, d.N, d.K, d.taps, d.groups = Mm, Nn, K, 1, 1
    d.A, d.lda, d.Wt, d.ldw = A.data_ptr(), A.shape[-1], W.data_ptr(), W.shape[-1]
    d.bias, d.bias_mod, d.act = N.ptr(bias), 1, act
    if resid is not None:
        d.resid, d.ldr = resid.data_ptr(), resid.shape[-1]
    d.rows_per_img = 1
    d.out, d.ldc = o.data_ptr(), o.shape[-1]
    return d
d1 = desc(M, H, Cc, x, w1, hid, bias=b1, act=2)
d3 = desc(M, Cc, H, hid, w3, out_a, resid=x)
two = lambda: (N.check(lib.sunb_gemm(C.byref(d1), 0, st), "conv1"), N.check(lib.sunb_gemm(C.byref(d3), 0, st), "conv3"))
fused = lambda: N.check(lib.sunb_mlp_fused(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w3.data_ptr(), out_b.data_ptr(), M, 0, 0, 0, st), "mlp_fused")

def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
two(); fused(); torch.cuda.synchronize()
n = min(M, 4096)
ref = x[:n].float() + torch.nn.functional.gelu(x[:n].float() @ w1.float().t() + b1).bfloat16().float() @ w3.float().t()
rel = lambda a, b: float((a.float() - b).norm() / b.norm())
print(f"rel err fused vs fp32 formula {rel(out_b[:n], ref):.2e}, two GEMMs {rel(out_a[:n], ref):.2e}, fused vs two GEMMs (all rows) {rel(out_b, out_a.float()):.2e}")
flops = 2 * 2.0 * M * Cc * H
tt, tf = t(two), t(fused)
print(f"B={B}: two GEMMs {tt:.1f} us ({flops/tt/1e6:.0f} TFLOP/s), fused {tf:.1f} us ({flops/tf/1e6:.0f} TFLOP/s)")

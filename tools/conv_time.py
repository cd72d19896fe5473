"""Time the stem conv2 / conv3 implicit GEMMs (2500 images) through sunb_gemm; env switches select the kernel variant."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "few-shot-vit_b200"))
from sunb200 import native as N
B = 2500
for (cin, cout) in ((64, 128), (128, 128)):
    a = torch.randn(B, 40, 40, cin, device="cuda").bfloat16()
    w = (torch.randn(9 * cout, cin, device="cuda") * (9 * cin) ** -0.5).bfloat16()
    bias = torch.randn(cout, device="cuda")
    out = torch.empty(B * 1600, cout, device="cuda", dtype=torch.bfloat16)
    d = N.GemmDesc()
    d.M, d.N, d.K, d.taps, d.groups = B * 1600, cout, cin, 9, 1
    d.a_mode, d.H, d.W, d.bw, d.bh = 1, 40, 40, 8, 8
    d.A, d.lda, d.Wt, d.ldw = a.data_ptr(), cin, w.data_ptr(), cin
    d.bias, d.bias_mod, d.act = bias.data_ptr(), 1, 1
    d.out, d.ldc = out.data_ptr(), cout
    st = N.current_stream()
    f = lambda: N.check(N.lib().sunb_gemm(C.byref(d), 0, st), "gemm")
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"conv {cin}->{cout}: {ms * 1000:.0f} us = {2.0 * B * 1600 * cin * cout * 9 / ms / 1e9:.0f} TFLOP/s", end="; ")
print("2CTA=" + os.environ.get("SUNB_CONV_SLAB_2CTA", "default"))

"""Debug/timing harness for sunb_gconv3x3 (run on the GPU box): error vs torch conv2d and time at the bench size."""
import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "few-shot-vit_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from sunb200 import native as N, packing
DEV = "cuda"
def run(B):
    g = torch.Generator(device="cpu").manual_seed(B)
    x = torch.randn(B * 400, 256, generator=g).to(DEV).bfloat16()
    w = (torch.randn(256, 32, 3, 3, generator=g) * (9 * 32) ** -0.5).to(DEV)
    wg = packing._grouped_taps(w).bfloat16().contiguous()
    y = torch.zeros(B * 400, 256, device=DEV, dtype=torch.bfloat16); y2 = torch.zeros_like(y)
    lib, st = N.lib(), N.current_stream()
    rc = lib.sunb_gconv3x3(x.data_ptr(), 256, wg.data_ptr(), y.data_ptr(), 256, y2.data_ptr(), 256, None, 0, B, 2, 0, st)
    torch.cuda.synchronize()
    if B <= 200:
        xf = x.float().view(B, 20, 20, 256).permute(0, 3, 1, 2)
        ref = F.conv2d(xf, w.bfloat16().float(), padding=1, groups=8).permute(0, 2, 3, 1).reshape(-1, 256)
        err = (y2.float() - ref).abs().max().item() / ref.abs().max().item()
        errg = (y.float() - F.gelu(ref)).abs().max().item() / ref.abs().max().item()
        print(f"B={B} rc={rc} rel err pre={err:.4e} gelu={errg:.4e}", flush=True)
        if err > 1e-2:
            d = (y2.float() - ref).abs().view(B, 400, 8, 32)
            print("  err by group", d.amax(dim=(0, 1, 3)).tolist())
            print("  err by image", d.amax(dim=(1, 2, 3)).tolist()[:8])
            print("  err by pixel row (img0)", d[0].view(20, 20, 8, 32).amax(dim=(1, 2, 3)).tolist())
    else:
        for _ in range(3):
            lib.sunb_gconv3x3(x.data_ptr(), 256, wg.data_ptr(), y.data_ptr(), 256, None, 0, None, 0, B, 2, 0, st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.sunb_gconv3x3(x.data_ptr(), 256, wg.data_ptr(), y.data_ptr(), 256, None, 0, None, 0, B, 2, 0, st)
        e1.record(); torch.cuda.synchronize()
        print(f"B={B} {e0.elapsed_time(e1) / 10 * 1000:.1f} us per launch (eval form: GELU, no y2)", flush=True)
for B in (1, 3, 77, 2500):
    run(B)

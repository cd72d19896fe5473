"""Time sunb_stem_in at 2500 images for the library named by SUNB200_LIB."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "few-shot-vit_b200"))
from sunb200 import native as N
B = 2500
x = torch.randn(B, 3, 80, 80, device="cuda")
w1, b1 = torch.randn(64, 27, device="cuda") * 0.2, torch.randn(64, device="cuda")
wd, bd = torch.randn(128, 27, device="cuda") * 0.2, torch.randn(128, device="cuda")
a1 = torch.empty(B * 1600, 64, device="cuda", dtype=torch.bfloat16)
idn = torch.empty(B * 1600, 128, device="cuda", dtype=torch.bfloat16)
st = N.current_stream()
f = lambda: N.check(N.lib().sunb_stem_in(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), wd.data_ptr(), bd.data_ptr(), a1.data_ptr(), idn.data_ptr(), B, 1, st), "stem_in")
for _ in range(3): f()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
print(f"stem_in: {e0.elapsed_time(e1) * 100:.0f} us  {os.environ.get('SUNB200_LIB', 'default')[-9:]}")

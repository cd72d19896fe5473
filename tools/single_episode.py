"""One 5-way 5-shot episode (100 images) through the public API, a few times (for an ncu launch list of the latency path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "few-shot-vit_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch, models, utils.few_shot as fs, sun_oracle as O
sd = O.calibrate_bn(O.init_meta_baseline_state_dict(12345))
m = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={})
m.load_state_dict(sd); m = m.cuda().eval()
x = torch.randn(100, 3, 80, 80, device="cuda")
xs, xq = fs.split_shot_query(x, 5, 5, 15)
with torch.no_grad():
    for _ in range(4):
        m(xs, xq)
torch.cuda.synchronize()

"""A/B timing of the eval forward (25 five-shot episodes = 2500 images) for the library named by SUNB200_LIB.
Run several times alternating builds inside ONE gpurun call; box-to-box variance is larger than most kernel changes."""
import os, sys, json, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "few-shot-vit_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import bench, models, sun_oracle as O, utils.few_shot as fs
dev = torch.device("cuda", 0)
sd = O.calibrate_bn(O.init_meta_baseline_state_dict(12345))
model = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={}); model.load_state_dict(sd); model = model.to(dev).eval()
x = bench.device_episodes(bench.CHUNK, 0, dev)
def fwd():
    xs, xq = fs.split_shot_query(x, bench.WAY, bench.SHOT, bench.QUERY, ep_per_batch=bench.CHUNK)
    return model(xs, xq)
with torch.no_grad():
    for _ in range(3): fwd()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5): fwd()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
print(f"{os.environ.get('SUNB200_LIB', 'default')}: {best:.3f} ms per 2500-image forward = {bench.CHUNK / best * 1000:.0f} eps/s")

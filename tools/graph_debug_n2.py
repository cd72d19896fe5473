"""2-GPU debug: which collective pattern breaks CUDA-graph capture (run under torchrun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "few-shot-vit_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist, torch.nn.functional as F
import models, utils, utils.few_shot as fs, sun_oracle as O
from sunb200.dist import GradAllReducer
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
sd = O.init_meta_baseline_state_dict(1)
data = torch.randn(2 * 3 * 3, 3, 80, 80, device=dev)
label = fs.make_nk_label(3, 2, 2).to(dev)
xs, xq = fs.split_shot_query(data, 3, 1, 2, ep_per_batch=2)
which = sys.argv[1]
m = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={}); m.load_state_dict(sd); m = m.to(dev).train()
opt, _ = utils.make_optimizer(m.parameters(), "sgd", lr=1e-3, weight_decay=5e-4)
if which == "overlap":
    m.encoder.enable_data_parallel(); red = GradAllReducer([m.temp])
else:
    red = GradAllReducer(m.parameters())
def f():
    loss = F.cross_entropy(m(xs, xq).view(-1, 3), label)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    if which != "none":
        red.all_reduce_mean()
    opt.step()
for _ in range(3): f()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        f()
    g.replay(); torch.cuda.synchronize()
    if rank == 0: print(f"{which}: capture OK")
except Exception as e:
    if rank == 0: print(f"{which}: FAILED {type(e).__name__}: {str(e).splitlines()[0]}")
dist.barrier(); dist.destroy_process_group()

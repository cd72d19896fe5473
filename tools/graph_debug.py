"""Find which part of the train step breaks CUDA-graph capture (debug helper, run on the GPU box)."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "few-shot-vit_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch, torch.nn.functional as F
import models, utils, utils.few_shot as fs, sun_oracle as O

dev = torch.device("cuda", 0)
sd = O.init_meta_baseline_state_dict(1)
data = torch.randn(2 * 3 * 3, 3, 80, 80, device=dev)
label = fs.make_nk_label(3, 2, 2).to(dev)

def build(rate):
    m = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={"drop_path_rate": rate})
    m.load_state_dict(sd)
    return m.to(dev)

def attempt(name, fn, mode):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(f"[{mode}] {name}: capture OK")
    except Exception as e:
        torch.cuda.synchronize()
        print(f"[{mode}] {name}: FAILED {type(e).__name__}: {str(e).splitlines()[0]}")

which = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "global"
xs, xq = fs.split_shot_query(data, 3, 1, 2, ep_per_batch=2)
if which == "eval":
    m = build(0.0).eval()
    with torch.no_grad():
        attempt("eval forward", lambda: m(xs, xq), mode)
elif which == "train_fwd":
    m = build(0.0).train()
    with torch.no_grad():
        pass
    def f():
        with torch.enable_grad():
            return m(xs, xq)
    attempt("train forward (rate 0)", f, mode)
elif which == "train_fwd_bwd":
    m = build(0.0).train()
    def f():
        loss = F.cross_entropy(m(xs, xq).view(-1, 3), label)
        m.zero_grad(set_to_none=True)
        loss.backward()
    attempt("train fwd+bwd (rate 0)", f, mode)
elif which == "train_dp":
    m = build(0.5).train()
    def f():
        loss = F.cross_entropy(m(xs, xq).view(-1, 3), label)
        m.zero_grad(set_to_none=True)
        loss.backward()
    attempt("train fwd+bwd (rate 0.5)", f, mode)
elif which == "train_full":
    m = build(0.5).train()
    opt, _ = utils.make_optimizer(m.parameters(), "sgd", lr=1e-3, weight_decay=5e-4)
    def f():
        loss = F.cross_entropy(m(xs, xq).view(-1, 3), label)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    attempt("train full step", f, mode)

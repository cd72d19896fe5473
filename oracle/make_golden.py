"""Generate tests/golden/*.npz by running the REAL reference modules (imported from /root/reference).

Run in the authoring container only (``python oracle/make_golden.py``): /root/reference does not exist
on the GPU box.  The reference modules are imported unmodified through the package stub of
SURVEY.md Appendix C; weights come from ``sun_oracle.init_*_state_dict`` (seeded, reproducible anywhere)
and are loaded with the reference's own ``load_state_dict``; inputs come from the seeded generators in
``sun_oracle``.  Only outputs (and checksums of the inputs) are stored, so the files stay small.

This script is test infrastructure; nothing in the product imports it.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import sun_oracle as O  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _load_pkg(sub, files, stub_timm=False):
    """Synthetic `models` package holding only the named reference files (Appendix C steps 1-3)."""
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]
    root = os.path.join(REF, sub)
    sys.path.insert(0, root)
    if stub_timm:
        timm = types.ModuleType("timm")
        tm = types.ModuleType("timm.models")
        tl = types.ModuleType("timm.models.layers")
        tl.DropPath = nn.Identity
        tl.to_2tuple = lambda x: (x, x)
        tl.trunc_normal_ = nn.init.trunc_normal_
        sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(root, "models")]
    sys.modules["models"] = pkg
    mods = {}
    for f in files:
        spec = importlib.util.spec_from_file_location(f"models.{f}", os.path.join(root, "models", f + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"models.{f}"] = m
        spec.loader.exec_module(m)
        mods[f] = m
        if f == "models":
            pkg.make, pkg.load, pkg.register = m.make, m.load, m.register
    import utils  # noqa: F401  (reference utils, torch only)
    import utils.few_shot  # noqa: F401
    sys.path.pop(0)
    return pkg, sys.modules["utils"]


def _checksum(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), (t * t).sum().item(), t[:: max(1, t.numel() // 97)].sum().item()])


def sd_checksum(sd) -> np.ndarray:
    return np.stack([_checksum(v) for _, v in sorted(sd.items())]).sum(0)


def hook_taps(encoder):
    """Forward hooks on the reference encoder at the layer boundaries the GPU tests compare."""
    taps = {}

    def keep(name):
        def fn(_m, _i, o):
            taps[name] = o.detach().clone()
        return fn
    hs = []
    for i, b in enumerate(encoder.stage1):
        hs.append(b.register_forward_hook(keep(f"stage1.{i}")))
    for i, b in enumerate(encoder.stage2):
        hs.append(b.register_forward_hook(keep(f"stage2.{i}")))
    for i, b in enumerate(encoder.stage3):
        hs.append(b.register_forward_hook(keep(f"stage3.{i}")))
    hs.append(encoder.stem.register_forward_hook(keep("stem_nopos")))
    hs.append(encoder.patch_embed2.register_forward_hook(keep("patch_embed2_nopos")))
    hs.append(encoder.patch_embed3.register_forward_hook(keep("patch_embed3_nopos")))
    hs.append(encoder.norm.register_forward_hook(keep("dense")))
    return taps, hs


def _grad_record(model, prefix, g):
    """Per-parameter gradient norms + values (small tensors) / strided samples (big tensors), as the GPU tests read them."""
    for k, p in model.named_parameters():
        gr = p.grad
        g[prefix + "gnorm." + k] = np.array(gr.norm().item())
        if gr.numel() <= 4096 or k in ("encoder.stem.conv1.weight", "encoder.stage3.2.attn.proj.weight"):
            g[prefix + "grad." + k] = gr.numpy().copy()
        else:
            g[prefix + "gsamp." + k] = gr.flatten()[:: max(1, gr.numel() // 2048)].numpy().copy()


def round2_fixtures():
    """Round-2 fixtures: SUN-M-shaped shards (10-way x (1+5) = 60 images per shard, drop_path 0.5, non-trivial loss) and
    their 2-shard gradient average; 20-episode argmax fixtures for 1-shot and 5-shot; one SUN meta-training step
    (offline.py:263-303, scaled down to 16 images -- stated) incl. an AdamW update."""
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    sd_w1 = O.calibrate_bn(O.init_meta_baseline_state_dict(12345))

    # ---------------- SUN-M shards (meta_tuning_sun_m/train_meta.py:161-174, configs/train_meta_mini_visformer_1shot.yaml:13-20)
    models, utils = _load_pkg("meta_tuning_sun_m", ["models", "visformer", "meta_baseline"])
    fs = sys.modules["utils.few_shot"]
    way, shot, query = 10, 1, 5
    g, shard_grads = {}, []
    for s in range(2):
        model = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={"drop_path_rate": 0.5})
        model.load_state_dict(sd_w1)
        model.train()
        data = O.make_episode_images(600 + s, way, shot + query, noise=1.0)
        xs, xq = fs.split_shot_query(data, way, shot, query, ep_per_batch=1)
        label = fs.make_nk_label(way, query, 1)
        torch.manual_seed(77 + s)                  # DropPath draws of this replica (global CPU RNG, visformer.py:94)
        logits = model(xs, xq).view(-1, way)
        loss = F.cross_entropy(logits, label)
        model.zero_grad()
        loss.backward()
        g[f"s{s}.x_checksum"] = _checksum(data)
        g[f"s{s}.logits"] = logits.detach().numpy()
        g[f"s{s}.loss"] = np.array(loss.item())
        shard_grads.append({k: p.grad.clone() for k, p in model.named_parameters()})
        if s == 0:
            _grad_record(model, "s0.", g)
            for k, v in model.state_dict().items():
                if k.endswith("running_mean") or k.endswith("running_var"):
                    g["s0.bn." + k] = v.numpy().copy()
        print("sunm shard", s, "loss", loss.item(), "acc", utils.compute_acc(logits, label))
    # DataParallel / DDP semantics: gradient of the batch mean = mean of the per-shard gradients
    for k, p in model.named_parameters():
        p.grad = (shard_grads[0][k] + shard_grads[1][k]) / 2
    _grad_record(model, "avg.", g)
    np.savez_compressed(os.path.join(OUT, "train_step_sunm.npz"), **g)

    # ---------------- argmax fixtures: 20 episodes each, BN-calibrated weights, class-structured inputs
    models, utils = _load_pkg("test_phase", ["models", "visformer", "meta_baseline"])
    fs = sys.modules["utils.few_shot"]
    model = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={})
    model.load_state_dict(sd_w1)
    model.eval()
    out = {}
    for tag, shot, seed0 in (("1shot", 1, 1000), ("5shot", 5, 2000)):
        lg = []
        for ep in range(20):
            data = O.make_episode_images(seed0 + ep, 5, shot + 15)
            xs, xq = fs.split_shot_query(data, 5, shot, 15, ep_per_batch=1)
            with torch.no_grad():
                lg.append(model(xs, xq)[0].numpy())
        out["logits_" + tag] = np.stack(lg)
        print("argmax fixture", tag, out["logits_" + tag].shape)
    np.savez_compressed(os.path.join(OUT, "episodes_argmax_w1.npz"), **out)

    # ---------------- SUN meta-training step (sun_meta_training/offline.py:263-303), batch scaled 512 -> 16 (stated)
    models, utils = _load_pkg("sun_meta_training", ["models", "visformer", "classifier", "token_label"], stub_timm=True)
    import models.visformer as vis_sun
    vis_sun.DEBUG = False
    src = open(os.path.join(REF, "sun_meta_training", "offline.py")).read().splitlines()
    ns = {"torch": torch, "nn": nn, "F": F}
    exec("\n".join(src[33:45]), ns)      # SoftTargetCrossEntropy, offline.py:34-45
    exec("\n".join(src[56:76]), ns)      # generate_softlabel, offline.py:57-76
    margs = dict(encoder="visformer_micro_80", encoder_args={"drop_path_rate": 0.5}, classifier="linear-classifier",
                 classifier_args={"n_classes": 64})
    student = models.make("token-label", **margs)
    teacher = models.make("token-label", **margs)
    sd_student = O.calibrate_bn(O.init_token_label_state_dict(4321))
    sd_teacher = O.calibrate_bn(O.init_token_label_state_dict(12345))
    student.load_state_dict(sd_student)
    teacher.load_state_dict(sd_teacher)
    teacher.eval()
    student.train()
    bs = 16
    strong = O.make_episode_images(800, 8, 2, noise=1.0)          # 8 classes x 2 images, "strong" view
    weak = O.make_episode_images(800, 8, 2, noise=0.5)            # same prototypes, "weak" view
    label = torch.arange(8).repeat_interleave(2) * 7 % 64
    lr = 5e-4 * (bs / 512)                                        # offline.py:228
    opt = torch.optim.AdamW(student.parameters(), betas=(0.9, 0.999), eps=1e-8, lr=lr, weight_decay=0.05)   # offline.py:229
    torch.manual_seed(123)
    logits_token, logits, token = student(strong)                 # offline.py:269
    cls_loss = F.cross_entropy(logits, label)
    with torch.no_grad():
        logits_token_t, _, _ = teacher(weak, True)                # offline.py:289
        soft_label = ns["generate_softlabel"](logits_token_t, k=5, bp=10, device="cpu")
    b, c, h, w = logits_token_t.size()
    token_loss = ns["SoftTargetCrossEntropy"]()(logits_token.permute(0, 2, 3, 1).view(-1, c + 1), soft_label)
    loss = cls_loss + 0.5 * token_loss                            # offline.py:300
    opt.zero_grad()
    loss.backward()
    g = {"loss": np.array(loss.item()), "cls_loss": np.array(cls_loss.item()), "token_loss": np.array(token_loss.item()),
         "soft_label": soft_label.numpy(), "teacher_logits_token": logits_token_t.numpy(), "logits": logits.detach().numpy(),
         "logits_token": logits_token.detach().numpy(), "label": label.numpy(), "lr": np.array(lr)}
    _grad_record(student, "", g)
    before = {k: p.detach().clone() for k, p in student.named_parameters()}
    opt.step()
    for k in ("classifier_local.linear.weight", "classifier.linear.bias", "encoder.stem.bn1.weight",
              "encoder.stage3.2.mlp.conv3.weight", "encoder.stage1.0.mlp.conv2.weight"):
        p = dict(student.named_parameters())[k]
        g["adamw_delta." + k] = (p.detach() - before[k]).flatten()[:: max(1, p.numel() // 1024)].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "sun_meta_step.npz"), **g)
    print("sun_meta_step loss", loss.item(), cls_loss.item(), token_loss.item())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)

    # ------------------------------------------------------------------ test_phase: eval path
    models, utils = _load_pkg("test_phase", ["models", "visformer", "meta_baseline"])
    fs = sys.modules["utils.few_shot"]

    model = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={})
    assert len(model.state_dict()) == 149
    names = [(k, tuple(v.shape)) for k, v in model.encoder.state_dict().items()]
    assert names == [(k, tuple(s)) for k, s in O.encoder_entries()], "oracle entry list != reference state_dict"

    # --- G1: encoder layer boundaries, randomized-BN weights (WR), 4 structured images, eval
    sd_wr = O.randomize_bn(O.init_meta_baseline_state_dict(12345), seed=7)
    model.load_state_dict(sd_wr)
    model.eval()
    x4 = O.make_episode_images(101, 2, 2)
    taps, hooks = hook_taps(model.encoder)
    with torch.no_grad():
        pooled = model.encoder(x4)
    for h in hooks:
        h.remove()
    g = {"x_checksum": _checksum(x4), "sd_checksum": sd_checksum(sd_wr), "pooled": pooled.numpy()}
    g["dense"] = taps["dense"].numpy()
    for k, v in taps.items():
        if k != "dense":
            g["tap0." + k] = v[0].numpy()          # image 0 only, keeps the file small
    np.savez_compressed(os.path.join(OUT, "encoder_eval_wr.npz"), **g)
    print("encoder_eval_wr", pooled.abs().mean().item())

    # --- G2: small episode through MetaBaseline (WR), 5-way 1-shot 3-query, eval
    data = O.make_episode_images(202, 5, 4)
    xs, xq = fs.split_shot_query(data, 5, 1, 3, ep_per_batch=1)
    with torch.no_grad():
        logits = model(xs, xq)
    label = fs.make_nk_label(5, 3, 1)
    np.savez_compressed(os.path.join(OUT, "episode_small_wr.npz"), x_checksum=_checksum(data),
                        logits=logits.numpy(), label=label.numpy(),
                        loss=F.cross_entropy(logits.view(-1, 5), label).item(),
                        acc=utils.compute_acc(logits.view(-1, 5), label))
    print("episode_small_wr", logits.flatten()[:5])

    # --- G3: full-size 5-way 1-shot 15-query episodes, BN-calibrated weights (W1), structured inputs (X1)
    sd_w1 = O.calibrate_bn(O.init_meta_baseline_state_dict(12345))
    model.load_state_dict(sd_w1)
    model.eval()
    outs, sums = [], []
    for ep in range(2):
        data = O.make_episode_images(300 + ep, 5, 16)
        xs, xq = fs.split_shot_query(data, 5, 1, 15, ep_per_batch=1)
        with torch.no_grad():
            outs.append(model(xs, xq)[0].numpy())
        sums.append(_checksum(data))
    np.savez_compressed(os.path.join(OUT, "episode_full_w1.npz"), logits=np.stack(outs), x_checksum=np.stack(sums),
                        sd_checksum=sd_checksum(sd_w1))
    print("episode_full_w1 logits range", np.stack(outs).min(), np.stack(outs).max())

    # --- G3b: raw-init weights (W0), iid inputs (X0): tolerance-only stress fixture
    sd_w0 = O.init_meta_baseline_state_dict(12345)
    model.load_state_dict(sd_w0)
    data = O.make_episode_images(400, 5, 16, structured=False)
    xs, xq = fs.split_shot_query(data, 5, 1, 15, ep_per_batch=1)
    with torch.no_grad():
        l0 = model(xs, xq)
    np.savez_compressed(os.path.join(OUT, "episode_full_w0.npz"), logits=l0.numpy(), x_checksum=_checksum(data))

    # --- G4: index utilities and compute_logits branches
    idx = {}
    for (way, shot, query, ep) in [(5, 1, 15, 1), (5, 5, 15, 2), (10, 1, 5, 8), (3, 2, 1, 4)]:
        n = ep * way * (shot + query)
        ids = torch.arange(n).view(n, 1, 1, 1).float().expand(n, 1, 1, 1).contiguous()
        s, q = fs.split_shot_query(ids, way, shot, query, ep_per_batch=ep)
        key = f"{way}_{shot}_{query}_{ep}"
        idx["shot_" + key] = s.flatten().long().numpy()
        idx["query_" + key] = q.flatten().long().numpy()
        idx["label_" + key] = fs.make_nk_label(way, query, ep).numpy()
    gl = torch.Generator().manual_seed(5)
    f3, p3 = torch.randn(2, 7, 16, generator=gl), torch.randn(2, 3, 16, generator=gl)
    for metric in ("dot", "cos", "sqr"):
        idx["logits3_" + metric] = utils.compute_logits(f3, p3, metric=metric, temp=2.5).numpy()
        idx["logits2_" + metric] = utils.compute_logits(f3[0], p3[0], metric=metric, temp=2.5).numpy()
    np.savez_compressed(os.path.join(OUT, "index_utils.npz"), **idx)

    # ------------------------------------------------------------------ meta_tuning_sun_m: train step
    models, utils = _load_pkg("meta_tuning_sun_m", ["models", "visformer", "meta_baseline"])
    fs = sys.modules["utils.few_shot"]
    for tag, rate in (("dp0", 0.0), ("dp05", 0.5)):
        model = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={"drop_path_rate": rate})
        model.load_state_dict(sd_w1)
        model.train()
        way, shot, query, ep = 3, 1, 2, 2
        data = O.make_episode_images(500, ep * way, shot + query)
        xs, xq = fs.split_shot_query(data, way, shot, query, ep_per_batch=ep)
        label = fs.make_nk_label(way, query, ep)
        torch.manual_seed(77)                 # DropPath draws come from the global RNG (visformer.py:94)
        logits = model(xs, xq).view(-1, way)
        loss = F.cross_entropy(logits, label)
        model.zero_grad()
        loss.backward()
        g = {"x_checksum": _checksum(data), "logits": logits.detach().numpy(), "loss": loss.item()}
        for k, p in model.named_parameters():
            gr = p.grad
            g["gnorm." + k] = np.array(gr.norm().item())
            if gr.numel() <= 4096 or k in ("encoder.stem.conv1.weight", "encoder.stage3.2.attn.proj.weight"):
                g["grad." + k] = gr.numpy()
            else:                               # a strided sample of the big tensors
                g["gsamp." + k] = gr.flatten()[:: max(1, gr.numel() // 2048)].numpy()
        for k, v in model.state_dict().items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                g["bn." + k] = v.numpy()
        # one SGD step as utils.make_optimizer builds it (lr 1e-3, momentum 0.9, wd 5e-4)
        opt, _ = utils.make_optimizer(model.parameters(), "sgd", lr=1e-3, weight_decay=5e-4)
        opt.step()
        g["after_sgd.temp"] = model.temp.detach().numpy()
        g["after_sgd.encoder.stem.bn1.weight"] = model.encoder.stem.bn1.weight.detach().numpy()
        g["after_sgd.encoder.stage3.2.attn.proj.weight.samp"] = \
            model.encoder.stage3[2].attn.proj.weight.detach().flatten()[::128].numpy()
        np.savez_compressed(os.path.join(OUT, f"train_step_{tag}.npz"), **g)
        print("train_step", tag, loss.item())

    # ------------------------------------------------------------------ sun_meta_training: SUN local head
    models, utils = _load_pkg("sun_meta_training", ["models", "visformer", "classifier", "token_label"], stub_timm=True)
    import models.visformer as vis_sun
    vis_sun.DEBUG = False    # the attention-matrix .cpu() clone is a debugging side effect, not an output
    student = models.make("token-label", encoder="visformer_micro_80", encoder_args={},
                          classifier="linear-classifier", classifier_args={"n_classes": 64})
    sd_tl = O.randomize_bn(O.init_token_label_state_dict(12345), seed=7)
    student.load_state_dict(sd_tl)
    student.eval()
    x4 = O.make_episode_images(101, 2, 2)
    with torch.no_grad():
        yt_s, y_s, tok_s = student(x4)
        yt_t, y_t, _ = student(x4, True)
    src = open(os.path.join(REF, "sun_meta_training", "offline.py")).read().splitlines()
    ns = {"torch": torch, "nn": nn, "F": F}
    exec("\n".join(src[33:45]), ns)      # SoftTargetCrossEntropy, offline.py:34-45
    exec("\n".join(src[56:76]), ns)      # generate_softlabel, offline.py:57-76
    gl = torch.Generator().manual_seed(11)
    t_logits = torch.randn(8, 5, 5, 64, generator=gl).permute(0, 3, 1, 2)      # NCHW view of NHWC memory
    soft = ns["generate_softlabel"](t_logits, k=5, bp=10, device="cpu")
    s_logits = torch.randn(8, 5, 5, 65, generator=gl).permute(0, 3, 1, 2)
    flat = s_logits.permute(0, 2, 3, 1).view(-1, 65)
    tl = ns["SoftTargetCrossEntropy"]()(flat, soft)
    soft_e2e = ns["generate_softlabel"](yt_t, k=5, bp=10, device="cpu")
    lab = torch.tensor([3, 3, 40, 40])
    total = F.cross_entropy(y_s, lab) + 0.5 * ns["SoftTargetCrossEntropy"]()(yt_s.permute(0, 2, 3, 1).view(-1, 65), soft_e2e)
    np.savez_compressed(os.path.join(OUT, "sun_head.npz"), sd_checksum=sd_checksum(sd_tl),
                        y_token_student=yt_s.numpy(), y_token_student_stride=np.array(yt_s.stride()),
                        y_student=y_s.numpy(), token_student=tok_s.numpy(),
                        y_token_teacher=yt_t.numpy(), soft_label=soft.numpy(), soft_ce=tl.item(),
                        soft_label_e2e=soft_e2e.numpy(), sun_loss=total.item())
    print("sun_head", tl.item(), total.item(), yt_s.stride())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    # `python oracle/make_golden.py`         -> round-1 fixtures (main) + round-2 fixtures
    # `python oracle/make_golden.py round2`  -> only the round-2 fixtures
    if "round2" not in sys.argv[1:]:
        main()
    round2_fixtures()

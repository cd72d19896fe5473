"""-m gpu: train-mode building blocks and the full meta-tuning step against torch autograd (fp32, same rounded inputs)
and the golden files produced by the real reference (tests/golden/train_step_*.npz).

Stated tolerances (bf16 operands, fp32 accumulate; BASELINE.md section 5 measures the reference itself under bf16
autocast at 4.8 % median / 13 % worst per-tensor gradient rel-L2 error):
  kernels vs fp32 torch on identical bf16 inputs : rel-L2 <= 1e-2
  full step vs fp32 reference                    : loss |delta| <= 0.05 * max(1, loss); per-tensor gradient rel-L2
                                                   <= 0.25 and cosine >= 0.97; running statistics rel <= 5e-2
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_helpers import rel_err, max_err, check_lib, check_call  # noqa: E402
from sunb200 import native as N, train as T, engine  # noqa: E402
import sun_oracle as O  # noqa: E402

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("P,Ma,Nb", [(1024, 128, 128), (1000, 256, 64), (777, 756, 256), (2000, 256, 252), (640, 65, 512)])
def test_wgrad_plain(P, Ma, Nb):
    lda, ldb = (Ma + 7) // 8 * 8, (Nb + 7) // 8 * 8
    dY = torch.full((P, lda), float("nan"), device=DEV, dtype=torch.bfloat16)
    X = torch.full((P, ldb), float("nan"), device=DEV, dtype=torch.bfloat16)
    dY[:, :Ma] = rnd(P, Ma, seed=1).bfloat16()
    X[:, :Nb] = rnd(P, Nb, seed=2).bfloat16()
    out = torch.zeros(Ma, Nb, device=DEV)
    T.wgrad(dY, X, out, P, Ma, Nb, Ca=Ma, Cb=Nb)
    torch.cuda.synchronize()
    ref = dY[:, :Ma].float().t() @ X[:, :Nb].float()
    assert rel_err(out, ref) < 2e-3
    T.wgrad(dY, X, out, P, Ma, Nb, Ca=Ma, Cb=Nb)          # accumulates
    torch.cuda.synchronize()
    assert rel_err(out, 2 * ref) < 2e-3


@pytest.mark.parametrize("B,H,Cin,Cout,box", [(3, 40, 64, 128, 8), (2, 40, 128, 128, 8), (5, 20, 128, 128, 4)])
def test_wgrad_conv(B, H, Cin, Cout, box):
    x = rnd(B, H, H, Cin, seed=3).bfloat16()
    dy = rnd(B, H, H, Cout, seed=4).bfloat16()
    out = torch.zeros(9 * Cout, Cin, device=DEV)
    T.wgrad(dy.view(-1, Cout), x.view(-1, Cin), out, B * H * H, Cout, Cin, taps=9, conv=(H, H, box, box))
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, 3, 3), dy.float().permute(0, 3, 1, 2),
                                      padding=1)
    got = out.view(9, Cout, Cin).permute(1, 2, 0).reshape(Cout, Cin, 3, 3)
    assert rel_err(got, ref) < 2e-3


def test_wgrad_grouped_and_dgrad():
    B, H = 3, 20
    x = rnd(B, H, H, 256, seed=5).bfloat16()
    dy = rnd(B, H, H, 256, seed=6).bfloat16()
    w = rnd(256, 32, 3, 3, seed=7, scale=0.06)
    scratch = torch.zeros(2 * 9 * 128, 128, device=DEV)
    T.wgrad(dy.view(-1, 256), x.view(-1, 256), scratch, B * H * H, 128, 128, Ca=256, Cb=256, taps=9, groups=2, a_goff=128,
            b_goff=128, conv=(H, H, 4, 4))
    dw = torch.zeros(256, 32, 3, 3, device=DEV)
    N.check(N.lib().sunb_grouped_wgrad_extract(scratch.data_ptr(), dw.data_ptr(), N.current_stream()), "extract")
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wf = w.bfloat16().float().requires_grad_(True)
    y = F.conv2d(xf, wf, padding=1, groups=8)
    y.backward(dy.float().permute(0, 3, 1, 2))
    assert rel_err(dw, wf.grad) < 2e-3
    # dgrad through the pair-packed, mirrored weights
    dst = torch.empty(4 * 9 * 64, 64, device=DEV, dtype=torch.bfloat16)
    N.check(N.lib().sunb_grouped_pairs(w.data_ptr(), dst.data_ptr(), 1, N.current_stream()), "pairs")
    dx = T.gemm(dy.view(-1, 256), dst, B * H * H, 64, 64, out=torch.empty(B * H * H, 256, device=DEV, dtype=torch.bfloat16),
                taps=9, groups=4, a_goff=64, c_goff=64, conv=(H, H, 4, 4))
    torch.cuda.synchronize()
    assert rel_err(dx, xf.grad.permute(0, 2, 3, 1).reshape(-1, 256)) < 1e-2


def test_bn_forward_backward_kernels():
    M, Cc = 5000, 256
    x = rnd(M, Cc, seed=8).bfloat16() * 2 + 0.5
    dz = rnd(M, Cc, seed=9).bfloat16()
    res = rnd(M, Cc, seed=10).bfloat16()
    gamma, beta = rnd(Cc, seed=11) * 0.2 + 1, rnd(Cc, seed=12) * 0.1
    P = {"bn.weight": gamma, "bn.bias": beta}
    Bf = {"bn.running_mean": torch.zeros(Cc, device=DEV), "bn.running_var": torch.ones(Cc, device=DEV),
          "bn.num_batches_tracked": torch.zeros((), dtype=torch.long, device=DEV)}
    eng = T.TrainEngine()
    eng.dev = torch.device(DEV)
    rec = eng.bn_forward(x, "bn", Cc, M, P, Bf)
    y = eng.bn_apply(x, rec, M, T.ACT_NONE)
    G = {"bn.weight": torch.zeros(Cc, device=DEV), "bn.bias": torch.zeros(Cc, device=DEV)}
    dx = eng.bn_backward(dz, rec, M, P, G, res=res)
    torch.cuda.synchronize()
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = torch.zeros(Cc, device=DEV), torch.ones(Cc, device=DEV)
    yr = F.batch_norm(xr, rm, rv, gr, br, training=True, momentum=0.1, eps=1e-5)
    yr.backward(dz.float())
    assert rel_err(y, yr) < 5e-3
    assert rel_err(dx, xr.grad + res.float()) < 1e-2
    assert rel_err(G["bn.weight"], gr.grad) < 2e-3 and rel_err(G["bn.bias"], br.grad) < 2e-3
    assert max_err(Bf["bn.running_mean"], rm) < 1e-4 and max_err(Bf["bn.running_var"], rv) < 1e-3
    assert int(Bf["bn.num_batches_tracked"]) == 1


@pytest.mark.parametrize("S,d,ds,B", [(100, 42, 42, 3), (25, 85, 85, 3)] +
                         [(S, d, ds, B) for S, d, ds in ((100, 42, 48), (25, 85, 96)) for B in (3, 7, 61)])
def test_attention_backward(S, d, ds, B):
    """Padded heads (ds = 48 / 96: the product's tcgen05 kernel, five images per tile for S = 25 -- B = 3, 7, 61 leave ragged
    last tiles) and the reference's packed layout (ds == d: the warp-MMA cross-check kernel of the test library) against
    autograd through the fp32 formula of visformer.py:183-190."""
    heads = 6
    inner = heads * ds
    ld3, ldi = (3 * inner + 15) // 16 * 16, (inner + 15) // 16 * 16
    val = torch.zeros(3 * heads, ds, device=DEV)
    val[:, :d] = 1.0                                                   # pad channels are exact zeros on input
    qkv = torch.zeros(B * S, ld3, device=DEV, dtype=torch.bfloat16)
    qkv[:, : 3 * inner] = (rnd(B * S, 3 * inner, seed=13) * val.reshape(1, -1)).bfloat16()
    dout = torch.zeros(B * S, ldi, device=DEV, dtype=torch.bfloat16)
    dout[:, :inner] = (rnd(B * S, inner, seed=14) * val[:heads].reshape(1, -1)).bfloat16()
    dqkv = torch.full((B * S, ld3), float("nan"), device=DEV, dtype=torch.bfloat16)
    if ds == d:
        check_call(check_lib().sunb_check_attention_backward(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, S, d, heads,
                                                             ld3, ldi, N.current_stream()), "check_attention_backward")
    else:
        N.check(N.lib().sunb_attention_backward(qkv.data_ptr(), dout.data_ptr(), dqkv.data_ptr(), B, S, d, ds, heads, ld3, ldi,
                                                N.current_stream()), "attention_backward")
    torch.cuda.synchronize()
    t = qkv[:, : 3 * inner].float().requires_grad_(True)
    u = t.reshape(B, S, 3, heads, ds).permute(2, 0, 3, 1, 4)
    p = torch.softmax(u[0] @ u[1].transpose(-1, -2) * d ** -0.5, dim=-1)
    o = (p @ u[2]).permute(0, 2, 1, 3).reshape(B * S, inner)
    o.backward(dout[:, :inner].float())
    got = dqkv[:, : 3 * inner].float()
    assert torch.isfinite(got).all()
    for x, nm in enumerate("qkv"):                                     # per-operand error: dq, dk, dv
        sl = slice(x * inner, (x + 1) * inner)
        assert rel_err(got[:, sl], t.grad[:, sl]) < 1e-2, nm
    if ds != d:                                                        # pad channels of the gradient are exact zeros
        assert float(got.reshape(B * S, 3 * heads, ds)[:, :, d:].abs().max()) == 0.0


@pytest.mark.parametrize("metric", ["cos", "sqr", "dot"])
def test_episode_logits_backward(metric):
    E, way, shot, Q, D = 2, 5, 3, 20, 512
    fs = rnd(E, way, shot, D, seed=15).requires_grad_(True)
    fq = rnd(E, Q, D, seed=16).requires_grad_(True)
    temp = torch.tensor(10.0, device=DEV, requires_grad=True)
    dl = rnd(E, Q, way, seed=17)
    out = engine.episode_logits(fs, fq, temp, metric)
    out.backward(dl)
    fs2, fq2 = fs.detach().cpu().requires_grad_(True), fq.detach().cpu().requires_grad_(True)
    t2 = torch.tensor(10.0, requires_grad=True)
    ref = O.compute_logits(fq2, fs2.mean(2), metric, t2)
    ref.backward(dl.cpu())
    assert rel_err(fs.grad.cpu(), fs2.grad) < 1e-4 and rel_err(fq.grad.cpu(), fq2.grad) < 1e-4
    assert abs(temp.grad.item() - t2.grad.item()) < 1e-3 * max(1.0, abs(t2.grad.item()))


def test_stem_wgrad_matches_conv2d_weight_grad():
    """sunb_stem_wgrad (im2col + tcgen05 split-K weight-gradient GEMMs) vs autograd of the two stride-2 3x3 stem convs."""
    B = 6
    x = rnd(B, 3, 80, 80, seed=70)
    da1 = rnd(B * 1600, 64, seed=71).bfloat16()
    didn = rnd(B * 1600, 128, seed=72).bfloat16()
    dw1 = torch.zeros(64, 27, device=DEV)
    dwd = torch.zeros(128, 27, device=DEV)
    scratch = torch.empty(B * 1600, 32, device=DEV, dtype=torch.bfloat16)
    lib = N.lib()
    N.check(lib.sunb_stem_wgrad(x.data_ptr(), da1.data_ptr(), didn.data_ptr(), dw1.data_ptr(), dwd.data_ptr(), B,
                                scratch.data_ptr(), N.current_stream()), "stem_wgrad")
    torch.cuda.synchronize()
    xb = x.bfloat16().float()                     # the kernel multiplies bf16-rounded pixels
    for g, dw, Cc in ((da1, dw1, 64), (didn, dwd, 128)):
        w = torch.zeros(Cc, 3, 3, 3, device=DEV, requires_grad=True)
        y = F.conv2d(xb, w, stride=2, padding=1)
        y.backward(g.float().view(B, 40, 40, Cc).permute(0, 3, 1, 2))
        assert rel_err(dw, w.grad.reshape(Cc, 27)) < 2e-3
    first = dw1.clone()
    # accumulates into the gradient buffers
    N.check(lib.sunb_stem_wgrad(x.data_ptr(), da1.data_ptr(), didn.data_ptr(), dw1.data_ptr(), dwd.data_ptr(), B,
                                scratch.data_ptr(), N.current_stream()), "stem_wgrad")
    torch.cuda.synchronize()
    assert rel_err(dw1, 2 * first) < 1e-5


def test_stem_tail_forward_backward():
    B = 3
    c3, idn = rnd(B * 1600, 128, seed=18).bfloat16(), rnd(B * 1600, 128, seed=19).bfloat16()
    s3, t3, sd, td = rnd(128, seed=20) * 0.2 + 1, rnd(128, seed=21) * 0.1, rnd(128, seed=22) * 0.2 + 1, rnd(128, seed=23) * 0.1
    pos = rnd(400, 128, seed=24) * 0.02
    g = rnd(B * 400, 128, seed=25).bfloat16()
    out = torch.empty(B * 400, 128, device=DEV, dtype=torch.bfloat16)
    dz = torch.empty(B * 1600, 128, device=DEV, dtype=torch.bfloat16)
    lib, st = N.lib(), N.current_stream()
    N.check(lib.sunb_stem_tail_forward(c3.data_ptr(), idn.data_ptr(), s3.data_ptr(), t3.data_ptr(), sd.data_ptr(),
                                       td.data_ptr(), pos.data_ptr(), out.data_ptr(), B, st), "tail fwd")
    N.check(lib.sunb_stem_tail_backward(c3.data_ptr(), idn.data_ptr(), s3.data_ptr(), t3.data_ptr(), sd.data_ptr(),
                                        td.data_ptr(), g.data_ptr(), dz.data_ptr(), B, st), "tail bwd")
    torch.cuda.synchronize()
    z = (c3.float() * s3 + t3 + idn.float() * sd + td).view(B, 40, 40, 128).permute(0, 3, 1, 2).requires_grad_(True)
    y = F.max_pool2d(F.leaky_relu(z, 0.1), 2).permute(0, 2, 3, 1).reshape(B * 400, 128) + pos.repeat(B, 1)
    y.backward(g.float())
    assert rel_err(out, y) < 5e-3
    assert rel_err(dz, z.grad.permute(0, 2, 3, 1).reshape(-1, 128)) < 5e-3


def _draw_dp_masks(seed, rate, batch):
    rates = O.drop_path_rates(rate)
    torch.manual_seed(seed)
    names = [f"stage1.{i}" for i in range(4)] + [f"stage2.{i}" for i in range(2)] + [f"stage3.{i}" for i in range(3)]
    masks = {}
    for bi, name in enumerate(names):
        if rates[bi] <= 0:
            continue
        n = 1 if name.startswith("stage1") else 2
        masks[name] = [torch.floor((1 - rates[bi]) + torch.rand(batch, 1, 1, 1)) for _ in range(n)]
    return masks, rates, names


@pytest.mark.parametrize("tag,rate", [("dp0", 0.0), ("dp05", 0.5)])
def test_meta_tuning_step_vs_reference(golden_dir, tag, rate):
    import models
    import utils
    import utils.few_shot as fs
    g = np.load(os.path.join(golden_dir, f"train_step_{tag}.npz"))
    sd = O.calibrate_bn(O.init_meta_baseline_state_dict(12345))
    model = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={"drop_path_rate": rate})
    model.load_state_dict(sd)
    model = model.cuda().train()
    way, shot, query, ep = 3, 1, 2, 2
    data = O.make_episode_images(500, ep * way, shot + query)
    xs, xq = fs.split_shot_query(data.cuda(), way, shot, query, ep_per_batch=ep)
    label = fs.make_nk_label(way, query, ep).cuda()
    scales = None
    if rate > 0:        # replay the reference's DropPath draws (CPU RNG stream under seed 77)
        masks, rates, names = _draw_dp_masks(77, rate, data.shape[0])
        scales = {n: [(m.view(-1) / (1 - rates[names.index(n)])).cuda() for m in ms] for n, ms in masks.items()}
        model.encoder._drop_path_scales = lambda batch, device: scales
    logits = model(xs, xq).view(-1, way)
    loss = F.cross_entropy(logits, label)
    model.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    print(f"[{tag}] loss {loss.item():.5f} vs reference {float(g['loss']):.5f}; "
          f"max |dlogit| {max_err(logits.detach().cpu(), torch.as_tensor(g['logits'])):.4f}")
    assert abs(loss.item() - float(g["loss"])) <= 0.05 * max(1.0, abs(float(g["loss"])))
    assert max_err(logits.detach().cpu(), torch.as_tensor(g["logits"])) < 0.35
    zero_grad = {"encoder.patch_embed2.proj.bias", "encoder.patch_embed2.norm.bn.bias",
                 "encoder.patch_embed3.proj.bias", "encoder.patch_embed3.norm.bn.bias"}
    report, bad = [], []
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        gr = p.grad.detach().cpu().float()
        ref_norm = float(g["gnorm." + name])
        if name in zero_grad:
            if gr.abs().max().item() > 2e-2 * max(1.0, ref_norm):
                bad.append((name, "nonzero", gr.abs().max().item()))
            continue
        if "grad." + name in g.files:
            ref = torch.as_tensor(g["grad." + name])
            rel = ((gr - ref).norm() / (ref.norm() + 1e-12)).item()
            cos = F.cosine_similarity(gr.flatten(), ref.flatten(), dim=0).item()
        else:
            ref = torch.as_tensor(g["gsamp." + name])
            samp = gr.flatten()[:: max(1, gr.numel() // 2048)]
            rel = ((samp - ref).norm() / (ref.norm() + 1e-12)).item()
            cos = F.cosine_similarity(samp, ref, dim=0).item()
        report.append((rel, cos, name))
        if not (rel <= 0.25 and cos >= 0.97):
            bad.append((name, rel, cos))
    report.sort(reverse=True)
    rels = sorted(r for r, _, _ in report)
    print(f"[{tag}] gradient rel-L2: median {rels[len(rels) // 2]:.4f}, worst {report[0][0]:.4f} ({report[0][2]})")
    for r, c, n in report[:8]:
        print(f"    {n:50s} rel {r:.4f} cos {c:.5f}")
    assert not bad, bad[:10]
    for k in g.files:
        if k.startswith("bn."):
            name = k[3:]
            got = model.state_dict()[name].cpu()
            ref = torch.as_tensor(g[k])
            assert ((got - ref).norm() / (ref.norm() + 1e-12)).item() < 5e-2, name
    # one SGD step exactly as the reference builds it, on the natively computed gradients
    opt, _ = utils.make_optimizer(model.parameters(), "sgd", lr=1e-3, weight_decay=5e-4)
    opt.step()
    assert abs(model.temp.item() - float(g["after_sgd.temp"])) < 1e-3
    got = model.encoder.stage3[2].attn.proj.weight.detach().flatten()[::128].cpu()
    assert rel_err(got, torch.as_tensor(g["after_sgd.encoder.stage3.2.attn.proj.weight.samp"])) < 1e-3


# ------------------------------------------------------------------------------------------------------------------
# SUN-M-shaped shards (BASELINE.json configs[2]: 10-way x (1 shot + 5 query) = 60 images per GPU at N = 8, drop_path 0.5,
# non-trivial loss) against tests/golden/train_step_sunm.npz, written by the real reference (oracle/make_golden.py).
# ------------------------------------------------------------------------------------------------------------------
ZERO_GRAD = {"encoder.patch_embed2.proj.bias", "encoder.patch_embed2.norm.bn.bias",
             "encoder.patch_embed3.proj.bias", "encoder.patch_embed3.norm.bn.bias"}


def compare_grads(grads, g, prefix, tag, rel_tol=0.25, cos_tol=0.97):
    """grads: {name: tensor}.  Same stated tolerance as the 18-image fixture: per-tensor rel-L2 <= 0.25, cosine >= 0.97;
    the four analytically-zero gradients are compared absolutely."""
    report, bad = [], []
    for name, gr in grads.items():
        gr = gr.detach().cpu().float()
        ref_norm = float(g[prefix + "gnorm." + name])
        if name in ZERO_GRAD:
            if gr.abs().max().item() > 2e-2 * max(1.0, ref_norm):
                bad.append((name, "nonzero", gr.abs().max().item()))
            continue
        if prefix + "grad." + name in g.files:
            ref = torch.as_tensor(g[prefix + "grad." + name])
            a = gr
        else:
            ref = torch.as_tensor(g[prefix + "gsamp." + name])
            a = gr.flatten()[:: max(1, gr.numel() // 2048)]
        rel = ((a - ref).norm() / (ref.norm() + 1e-12)).item()
        cos = F.cosine_similarity(a.flatten(), ref.flatten(), dim=0).item()
        report.append((rel, cos, name))
        if not (rel <= rel_tol and cos >= cos_tol):
            bad.append((name, rel, cos))
    report.sort(reverse=True)
    rels = sorted(r for r, _, _ in report)
    print(f"[{tag}] gradient rel-L2 over {len(rels)} tensors: median {rels[len(rels) // 2]:.4f}, worst {report[0][0]:.4f} ({report[0][2]})")
    for r, c, n in report[:5]:
        print(f"    {n:50s} rel {r:.4f} cos {c:.5f}")
    assert not bad, bad[:10]
    return rels[len(rels) // 2], report[0][0]


def sunm_shard_grads(shard, device="cuda"):
    """One replica's step on SUN-M shard `shard`: fresh model from the calibrated weights, the reference's DropPath draws
    replayed (CPU RNG stream under seed 77 + shard).  Returns (model, logits, loss)."""
    import models
    import utils.few_shot as fs
    way, shot, query = 10, 1, 5
    sd = O.calibrate_bn(O.init_meta_baseline_state_dict(12345))
    model = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={"drop_path_rate": 0.5})
    model.load_state_dict(sd)
    model = model.to(device).train()
    data = O.make_episode_images(600 + shard, way, shot + query, noise=1.0).to(device)
    xs, xq = fs.split_shot_query(data, way, shot, query, ep_per_batch=1)
    label = fs.make_nk_label(way, query, 1).to(device)
    masks, rates, names = _draw_dp_masks(77 + shard, 0.5, data.shape[0])
    scales = {n: [(m.view(-1) / (1 - rates[names.index(n)])).to(device) for m in ms] for n, ms in masks.items()}
    model.encoder._drop_path_scales = lambda batch, dev: scales
    logits = model(xs, xq).view(-1, way)
    loss = F.cross_entropy(logits, label)
    model.zero_grad()
    return model, logits, loss


def test_sunm_shard_step_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_step_sunm.npz"))
    model, logits, loss = sunm_shard_grads(0)
    loss.backward()
    torch.cuda.synchronize()
    print(f"[sunm s0] loss {loss.item():.4f} vs reference {float(g['s0.loss']):.4f}; "
          f"max |dlogit| {max_err(logits.detach().cpu(), torch.as_tensor(g['s0.logits'])):.4f}")
    assert abs(loss.item() - float(g["s0.loss"])) <= 0.05 * float(g["s0.loss"])
    assert max_err(logits.detach().cpu(), torch.as_tensor(g["s0.logits"])) < 0.35
    compare_grads({n: p.grad for n, p in model.named_parameters()}, g, "s0.", "sunm shard 0")
    for k in g.files:
        if k.startswith("s0.bn."):
            name = k[6:]
            got, ref = model.state_dict()[name].cpu(), torch.as_tensor(g[k])
            assert ((got - ref).norm() / (ref.norm() + 1e-12)).item() < 5e-2, name


def test_sunm_two_shards_sequential_average_vs_reference(golden_dir):
    """Data-parallel semantics on ONE GPU: the two shards run one after the other as two replicas would, their gradients
    are averaged, and the average is compared with the reference's 2-shard average (what N = 2 ranks must produce)."""
    g = np.load(os.path.join(golden_dir, "train_step_sunm.npz"))
    acc = None
    for s in range(2):
        model, logits, loss = sunm_shard_grads(s)
        loss.backward()
        torch.cuda.synchronize()
        assert abs(loss.item() - float(g[f"s{s}.loss"])) <= 0.05 * float(g[f"s{s}.loss"])
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
        acc = grads if acc is None else {n: (acc[n] + grads[n]) / 2 for n in acc}
    compare_grads(acc, g, "avg.", "sunm 2-shard average")


# ------------------------------------------------------------------------------------------------------------------
# fused optimizers and the SUN meta-training step (sun_meta_training/offline.py:263-303)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["sgd", "adamw"])
def test_fused_optimizers_match_torch(kind):
    from sunb200.optim import FusedSGD, FusedAdamW
    shapes = [(64, 3, 3, 3), (128,), (), (1530, 512, 1, 1), (65, 512), (7,), (4097,)]
    ps = [torch.nn.Parameter(rnd(*s, seed=100 + i) if len(s) else rnd(1, seed=100 + i)[0].clone()) for i, s in enumerate(shapes)]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    if kind == "sgd":
        a, b = FusedSGD(ps, lr=1e-2, momentum=0.9, weight_decay=5e-4), torch.optim.SGD(qs, lr=1e-2, momentum=0.9, weight_decay=5e-4)
    else:
        a = FusedAdamW(ps, lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
        b = torch.optim.AdamW(qs, lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    for it in range(4):
        for i, (p, q) in enumerate(zip(ps, qs)):
            gr = rnd(*p.shape, seed=1000 * it + i) if p.dim() else rnd(1, seed=1000 * it + i)[0].clone()
            p.grad, q.grad = gr.clone(), gr.clone()
        if it == 2:                                   # scheduler changes the learning rate between steps
            for o in (a, b):
                o.param_groups[0]["lr"] *= 0.5
        a.step()
        b.step()
    torch.cuda.synchronize()
    for p, q, s in zip(ps, qs, shapes):
        assert (p - q).abs().max().item() <= 2e-6 * max(1.0, q.abs().max().item()), s


def test_sun_meta_training_step_vs_reference(golden_dir):
    """One SUN meta-training step (offline.py:263-303, batch scaled 512 -> 16): losses, gradients of all 90 tensors and the
    AdamW update against the reference's own outputs (tests/golden/sun_meta_step.npz)."""
    import models
    from sunb200 import sun_meta
    g = np.load(os.path.join(golden_dir, "sun_meta_step.npz"))
    margs = dict(encoder="visformer_micro_80", encoder_args={"drop_path_rate": 0.5}, classifier="linear-classifier",
                 classifier_args={"n_classes": 64})
    student, teacher = models.make("token-label", **margs), models.make("token-label", **margs)
    student.load_state_dict(O.calibrate_bn(O.init_token_label_state_dict(4321)))
    teacher.load_state_dict(O.calibrate_bn(O.init_token_label_state_dict(12345)))
    student, teacher = student.cuda().train(), teacher.cuda().eval()
    strong = O.make_episode_images(800, 8, 2, noise=1.0).cuda()
    weak = O.make_episode_images(800, 8, 2, noise=0.5).cuda()
    label = torch.as_tensor(g["label"]).cuda()
    masks, rates, names = _draw_dp_masks(123, 0.5, 16)
    scales = {n: [(m.view(-1) / (1 - rates[names.index(n)])).cuda() for m in ms] for n, ms in masks.items()}
    student.encoder._drop_path_scales = lambda batch, dev: scales
    # (1) end to end: our bf16 teacher's pseudo labels vs the fp32 reference's (reported statistic)
    with torch.no_grad():
        yt_t, _, _ = teacher(weak, True)
        soft = engine.generate_softlabel(yt_t, k=5, bp=10)
    ref_soft = torch.as_tensor(g["soft_label"]).cuda()
    same_rows = (soft == ref_soft).all(dim=1).float().mean().item()
    print(f"[sun step] teacher patch logits rel-L2 {rel_err(yt_t.cpu(), torch.as_tensor(g['teacher_logits_token'])):.4f}; "
          f"soft-label rows identical to the fp32 reference: {100 * same_rows:.1f} %")
    assert rel_err(yt_t.cpu(), torch.as_tensor(g["teacher_logits_token"])) < 5e-2
    assert same_rows > 0.7
    # (2) the step itself with the reference's soft labels injected (identical targets -> comparable gradients)
    before = {k: p.detach().clone() for k, p in student.named_parameters()}
    opt, _ = sun_meta.build_optimizer(student, batch_size=16)
    assert abs(opt.param_groups[0]["lr"] - 1e-6) < 1e-12           # warm-up start of the cosine schedule
    opt.param_groups[0]["lr"] = float(g["lr"])
    out = sun_meta.sun_meta_training_step(student, teacher, strong, weak, label, opt, 5, 10, soft_label=ref_soft)
    torch.cuda.synchronize()
    print(f"[sun step] loss {out['loss'].item():.4f} / cls {out['cls_loss'].item():.4f} / token {out['token_loss'].item():.4f}"
          f" vs reference {float(g['loss']):.4f} / {float(g['cls_loss']):.4f} / {float(g['token_loss']):.4f}")
    for k in ("loss", "cls_loss", "token_loss"):
        assert abs(out[k].item() - float(g[k])) <= 0.05 * float(g[k]), k
    compare_grads({n: p.grad for n, p in student.named_parameters()}, g, "", "sun meta step")
    # AdamW's first step moves every element by ~lr * sign(g): compare the update direction and size
    for k in g.files:
        if k.startswith("adamw_delta."):
            name = k[len("adamw_delta."):]
            p = dict(student.named_parameters())[name]
            d = (p.detach() - before[name]).flatten()[:: max(1, p.numel() // 1024)].cpu()
            ref = torch.as_tensor(g[k])
            agree = (torch.sign(d) == torch.sign(ref)).float().mean().item()
            assert agree > 0.9 and abs(d.abs().mean().item() / ref.abs().mean().item() - 1) < 0.1, (name, agree)

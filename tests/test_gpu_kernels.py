"""-m gpu: every kernel behind the C ABI against a torch fp32 reference of the same op on the same bf16-rounded
inputs (tolerances cover only fp32 summation order and the final bf16 rounding of the output)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gpu_helpers import run_gemm, act_ref, rel_err, max_err, check_lib, check_call  # noqa: E402
from sunb200 import native as N, engine  # noqa: E402
import sun_oracle as O  # noqa: E402

DEV = "cuda"
BF16_OUT = 6e-3      # relative L2 tolerance for bf16 outputs
F32_OUT = 2e-4


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("M,Nn,K,act,use_bias,use_res", [
    (256, 128, 64, 0, False, False),
    (300, 256, 128, 2, True, False),        # ragged M, GELU, bias
    (1000, 756, 256, 0, True, False),       # qkv: ragged N
    (1000, 256, 252, 0, False, True),       # proj: K tail (TMA zero fill), residual
    (500, 512, 510, 0, False, True),
    (2000, 1024, 256, 2, True, False),
    (2000, 256, 1024, 0, False, True),
    (128, 65, 512, 0, True, False),         # SUN local classifier shape (odd N)
])
def test_gemm_plain(impl, M, Nn, K, act, use_bias, use_res):
    lda = (K + 7) // 8 * 8
    A = torch.zeros(M, lda, device=DEV, dtype=torch.bfloat16)
    A[:, :K] = rnd(M, K, seed=1).bfloat16()
    A[:, K:] = float("nan")                      # padding columns must never be read
    Wt = torch.zeros(Nn, lda, device=DEV, dtype=torch.bfloat16)
    Wt[:, :K] = rnd(Nn, K, seed=2, scale=K ** -0.5).bfloat16()
    Wt[:, K:] = float("nan")
    bias = rnd(Nn, seed=3) if use_bias else None
    ldc = (Nn + 7) // 8 * 8
    resid = rnd(M, ldc, seed=4).bfloat16() if use_res else None
    out = run_gemm(A, Wt, M, Nn, K, impl=impl, bias=bias, act=act, resid=resid, out_cols=ldc)
    ref = A[:, :K].float() @ Wt[:, :K].float().t()
    if use_res:
        ref = ref + resid[:, :Nn].float()
    if use_bias:
        ref = ref + bias
    ref = act_ref(ref, act)
    assert rel_err(out[:, :Nn], ref) < BF16_OUT
    if ldc > Nn:
        assert (out[:, Nn:] == 0).all()          # columns past N untouched


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_gemm_f32_out_and_bias_table(impl):
    M, Nn, K = 400, 256, 512
    A = rnd(M, K, seed=5).bfloat16()
    Wt = rnd(Nn, K, seed=6, scale=K ** -0.5).bfloat16()
    table = rnd(100, Nn, seed=7)
    out = run_gemm(A, Wt, M, Nn, K, impl=impl, bias=table, bias_mod=100, out_f32=True)
    ref = A.float() @ Wt.float().t() + table.repeat(4, 1)
    assert rel_err(out, ref) < F32_OUT


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_gemm_row_scale_and_s2d(impl):
    B, H, Wd, Cc, K = 3, 10, 10, 256, 128
    M = B * H * Wd
    A = rnd(M, K, seed=8).bfloat16()
    Wt = rnd(Cc, K, seed=9, scale=K ** -0.5).bfloat16()
    resid = rnd(M, Cc, seed=10).bfloat16()
    rs = torch.tensor([0.0, 2.0, 1.0], device=DEV)
    out = run_gemm(A, Wt, M, Cc, K, impl=impl, resid=resid, row_scale=rs, rows_per_img=H * Wd, out_map=1, oHW=(H, Wd))
    ref = (A.float() @ Wt.float().t()) * rs.repeat_interleave(H * Wd)[:, None] + resid.float()
    ref = ref.reshape(B, H // 2, 2, Wd // 2, 2, Cc).permute(0, 1, 3, 2, 4, 5).reshape(M, Cc)
    assert rel_err(out, ref) < BF16_OUT


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("B,H,Cin,Cout,box", [(3, 40, 64, 128, 8), (2, 40, 128, 128, 8), (5, 20, 64, 64, 4)])
def test_conv3x3(impl, B, H, Cin, Cout, box):
    x = rnd(B, H, H, Cin, seed=11).bfloat16()
    w = rnd(Cout, Cin, 3, 3, seed=12, scale=(9 * Cin) ** -0.5)
    taps = w.permute(2, 3, 0, 1).reshape(9 * Cout, Cin).bfloat16().contiguous()
    bias = rnd(Cout, seed=13)
    resid = rnd(B * H * H, Cout, seed=14).bfloat16()
    out = run_gemm(x, taps, B * H * H, Cout, Cin, impl=impl, taps=9, conv=(H, H, box, box), bias=bias, act=1, resid=resid)
    wq = taps.float().reshape(3, 3, Cout, Cin).permute(2, 3, 0, 1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wq, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    ref = F.leaky_relu(ref + resid.float() + bias, 0.1)
    assert rel_err(out, ref) < BF16_OUT


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_grouped_conv_pairs(impl):
    from sunb200 import packing
    B, H = 3, 20
    x = rnd(B, H, H, 256, seed=15).bfloat16()
    w = rnd(256, 32, 3, 3, seed=16, scale=(9 * 32) ** -0.5)
    pairs = packing._grouped_pairs(w).bfloat16().reshape(4 * 9 * 64, 64).contiguous()
    out = run_gemm(x, pairs, B * H * H, 64, 64, impl=impl, taps=9, groups=4, a_goff=64, c_goff=64,
                   conv=(H, H, 4, 4), act=2)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), padding=1, groups=8)
    ref = act_ref(ref, 2).permute(0, 2, 3, 1).reshape(-1, 256)
    assert rel_err(out, ref) < BF16_OUT


@pytest.mark.parametrize("S,d,C", [(100, 42, 256), (25, 85, 512)])
def test_attention_checker_packed_layout(S, d, C):
    """The warp-MMA cross-check kernel (test library) on the reference's packed head layout vs the fp32 formula."""
    B, heads = 5, 6
    inner = heads * d
    ld_qkv = (3 * inner + 7) // 8 * 8
    qkv = torch.full((B * S, ld_qkv), float("nan"), device=DEV, dtype=torch.bfloat16)
    qkv[:, : 3 * inner] = rnd(B * S, 3 * inner, seed=17).bfloat16()
    out = torch.zeros(B * S, C, device=DEV, dtype=torch.bfloat16)
    check_call(check_lib().sunb_check_attention(qkv.data_ptr(), out.data_ptr(), B, S, d, d, heads, ld_qkv, C, N.current_stream()),
               "check_attention")
    torch.cuda.synchronize()
    t = qkv[:, : 3 * inner].float().reshape(B, S, 3, heads, d).permute(2, 0, 3, 1, 4)
    p = torch.softmax(t[0] @ t[1].transpose(-1, -2) * d ** -0.5, dim=-1)
    ref = (p @ t[2]).permute(0, 2, 1, 3).reshape(B * S, inner)
    assert rel_err(out[:, :inner], ref) < BF16_OUT
    assert (out[:, inner:] == 0).all()


def test_attention_rejects_unsupported_layouts():
    """The product has ONE attention implementation (tcgen05, padded heads); other layouts fail loudly instead of falling back."""
    B, S, d, heads = 2, 100, 42, 6
    qkv = torch.zeros(B * S, 768, device=DEV, dtype=torch.bfloat16)
    out = torch.zeros(B * S, 256, device=DEV, dtype=torch.bfloat16)
    assert N.lib().sunb_attention(qkv.data_ptr(), out.data_ptr(), B, S, d, d, heads, 768, 256, N.current_stream()) != 0
    assert b"unsupported" in N.lib().sunb_last_error()
    assert N.lib().sunb_attention_backward(qkv.data_ptr(), out.data_ptr(), qkv.data_ptr(), B, S, d, d, heads, 768, 256,
                                           N.current_stream()) != 0
    assert b"unsupported" in N.lib().sunb_last_error()


@pytest.mark.parametrize("B", [7, 301])
@pytest.mark.parametrize("S,d,dp", [(100, 42, 48), (25, 85, 96)])
def test_attention_padded_heads(S, d, dp, B):
    """Eval-engine layout: heads padded to dp channels (zeros in, zeros out): the tcgen05 kernel (attention_tc.cu).
    B = 7 exercises a partial 5-image tile (S = 25), B = 301 the multi-tile software pipeline of every persistent CTA."""
    heads = 6
    real = rnd(B * S, 3, heads, d, seed=21).bfloat16()
    qkv = torch.zeros(B * S, 3, heads, dp, device=DEV, dtype=torch.bfloat16)
    qkv[..., :d] = real
    qkv = qkv.reshape(B * S, 3 * heads * dp).contiguous()
    out = torch.full((B * S, heads * dp), float("nan"), device=DEV, dtype=torch.bfloat16)
    N.check(N.lib().sunb_attention(qkv.data_ptr(), out.data_ptr(), B, S, d, dp, heads, 3 * heads * dp, heads * dp,
                                   N.current_stream()), "attention")
    torch.cuda.synchronize()
    t = real.float().reshape(B, S, 3, heads, d).permute(2, 0, 3, 1, 4)
    p = torch.softmax(t[0] @ t[1].transpose(-1, -2) * d ** -0.5, dim=-1)
    ref = (p @ t[2]).permute(0, 2, 1, 3)                      # [B, S, heads, d]
    o = out.reshape(B, S, heads, dp)
    assert rel_err(o[..., :d], ref) < BF16_OUT
    assert (o[..., d:] == 0).all()                            # pad channels written as exact zeros
    # second opinion: the warp-MMA checker on the same padded buffers
    out2 = torch.full_like(out, float("nan"))
    check_call(check_lib().sunb_check_attention(qkv.data_ptr(), out2.data_ptr(), B, S, d, dp, heads, 3 * heads * dp, heads * dp,
                                                N.current_stream()), "check_attention")
    torch.cuda.synchronize()
    assert rel_err(o[..., :d], out2.reshape(B, S, heads, dp)[..., :d]) < BF16_OUT


@pytest.mark.parametrize("metric", ["cos", "dot", "sqr"])
def test_episode_logits(metric):
    E, way, shot, Q, D = 3, 5, 5, 75, 512
    fs, fq = rnd(E, way, shot, D, seed=18), rnd(E, Q, D, seed=19)
    temp = torch.tensor(10.0, device=DEV)
    out = engine.episode_logits(fs, fq, temp, metric)
    ref = O.compute_logits(fq.cpu(), fs.cpu().mean(2), metric, 10.0)
    assert max_err(out.cpu(), ref) < 2e-4 * max(1.0, ref.abs().max().item())


def test_ce_and_acc():
    logits, label = rnd(75, 5, seed=20), torch.randint(0, 5, (75,), generator=torch.Generator().manual_seed(1)).to(DEV)
    out = engine.ce_and_acc(logits, label).cpu()
    assert abs(out[0].item() - F.cross_entropy(logits, label).item()) < 1e-5
    assert out[1].item() == pytest.approx((logits.argmax(1) == label).float().mean().item())


def test_softlabel_bit_exact(golden_dir):
    import numpy as np, os
    g = np.load(os.path.join(golden_dir, "sun_head.npz"))
    gl = torch.Generator().manual_seed(11)
    t_logits = torch.randn(8, 5, 5, 64, generator=gl).to(DEV).permute(0, 3, 1, 2)     # NHWC-backed NCHW view
    soft = engine.generate_softlabel(t_logits, k=5, bp=10)
    assert np.array_equal(soft.cpu().numpy(), g["soft_label"])
    soft2 = engine.generate_softlabel(t_logits.contiguous(), k=5, bp=10)                # NCHW-contiguous input
    assert np.array_equal(soft2.cpu().numpy(), g["soft_label"])
    for (k, bp) in [(3, 10), (1, 0), (5, 25), (64, 3)]:
        ref = O.generate_softlabel(t_logits.cpu(), k=k, bp=bp)
        assert torch.equal(engine.generate_softlabel(t_logits, k=k, bp=bp).cpu(), ref), (k, bp)


def test_softlabel_many_classes():
    """tieredImageNet has 351 base classes: class chunks per lane cover n_cls <= 512."""
    gl = torch.Generator().manual_seed(12)
    t_logits = torch.randn(3, 5, 5, 351, generator=gl).to(DEV).permute(0, 3, 1, 2)
    ref = O.generate_softlabel(t_logits.cpu(), k=5, bp=10)
    assert torch.equal(engine.generate_softlabel(t_logits, k=5, bp=10).cpu(), ref)


@pytest.mark.parametrize("lead,K,n", [((96,), 512, 64), ((8, 5, 5), 512, 65), ((33,), 256, 10)])
def test_linear_forward_backward(lead, K, n):
    """utils.linear: native forward, dgrad (sunb_gemm), wgrad (sunb_wgrad) and bias gradient vs F.linear on the same
    bf16-rounded operands (reference heads: sun_meta_training/models/classifier.py:27-35, token_label.py:48-60)."""
    import utils
    x = rnd(*lead, K, seed=60).bfloat16().float().requires_grad_(True)
    w = (rnd(n, K, seed=61) * K ** -0.5).bfloat16().float().requires_grad_(True)
    b = rnd(n, seed=62).requires_grad_(True)
    dy = rnd(*lead, n, seed=63).bfloat16().float()
    y = utils.linear(x, w, b)
    assert y.grad_fn is not None and y.shape == (*lead, n)
    y.backward(dy)
    torch.cuda.synchronize()
    xr, wr, br = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.linear(xr, wr, br)
    yr.backward(dy)
    assert rel_err(y, yr) < 2e-3
    assert rel_err(x.grad, xr.grad) < 1e-2            # dY and W are rounded to bf16 for the tensor-core dgrad
    assert rel_err(w.grad, wr.grad) < 1e-2
    assert rel_err(b.grad, br.grad) < 1e-2
    # a bf16 copy of the activations can be handed in (what the encoder's last kernel emits)
    y2 = utils.linear(x.detach(), w.detach(), b.detach(), x_bf16=x.detach().bfloat16())
    assert torch.equal(y2, y.detach())


def test_soft_ce_forward_backward(golden_dir):
    import numpy as np, os
    g = np.load(os.path.join(golden_dir, "sun_head.npz"))
    gl = torch.Generator().manual_seed(11)
    torch.randn(8, 5, 5, 64, generator=gl)
    s_logits = torch.randn(8, 5, 5, 65, generator=gl).reshape(-1, 65)
    soft = torch.as_tensor(g["soft_label"])
    x = s_logits.to(DEV).requires_grad_(True)
    loss = engine.soft_target_cross_entropy(x, soft.to(DEV))
    assert abs(loss.item() - float(g["soft_ce"])) < 1e-4
    (loss * 0.5).backward()
    xr = s_logits.clone().requires_grad_(True)
    (O.soft_target_cross_entropy(xr, soft) * 0.5).backward()
    assert max_err(x.grad.cpu(), xr.grad) < 1e-6
    # tiled target (rows of x an integer multiple of target rows, offline.py:41-43)
    x2 = torch.cat([s_logits, s_logits * 0.5]).to(DEV)
    l2 = engine.soft_target_cross_entropy(x2, soft.to(DEV))
    assert abs(l2.item() - O.soft_target_cross_entropy(x2.cpu(), soft).item()) < 1e-4


def test_layernorm_wrapper():
    from models.visformer import LayerNorm
    ln = LayerNorm(96).cuda()
    with torch.no_grad():
        ln.ln.weight.copy_(rnd(96, seed=30) * 0.2 + 1)
        ln.ln.bias.copy_(rnd(96, seed=31) * 0.1)
    x = rnd(3, 96, 7, 5, seed=32) * 3 + 1
    ref = F.layer_norm(x.permute(0, 2, 3, 1), (96,), ln.ln.weight, ln.ln.bias, ln.ln.eps).permute(0, 3, 1, 2)
    assert max_err(ln(x), ref) < 1e-5


def test_gelu_device_accuracy():
    """The device GELU (tanh-form fit, common.cuh gelu_fast) and its training derivative (erf form) against exact erf
    GELU in fp64: identity GEMM -> fp32 output, so the only error is the activation's own."""
    x = torch.linspace(-12, 12, 64 * 4096, device=DEV).bfloat16().view(-1, 64).contiguous()
    eye = torch.eye(64, device=DEV).bfloat16().contiguous()
    out = run_gemm(x, eye, x.shape[0], 64, 64, act=2, out_f32=True)
    xd = x.double()
    ref = 0.5 * xd * (1 + torch.erf(xd * 0.70710678118654752))
    err = (out.double() - ref).abs()
    # absolute error bound everywhere, and relative to the bf16 rounding step of the value wherever |gelu| >= 0.05
    assert err.max().item() < 1.5e-3, err.max().item()
    big = ref.abs() >= 0.05
    assert (err[big] / ref[big].abs()).max().item() < 2e-3          # bf16 half-ulp is 3.9e-3
    assert err[xd.abs() <= 2].max().item() < 4e-4
    # derivative through the chain-rule epilogue: out = 1 * gelu'(aux)
    ones = torch.zeros_like(x); ones[:, 0] = 1
    w1 = torch.zeros(64, 64, device=DEV).bfloat16(); w1[:, 0] = 1
    d = N.GemmDesc()
    d.M, d.N, d.K, d.taps, d.groups = x.shape[0], 64, 64, 1, 1
    d.A, d.lda, d.Wt, d.ldw = ones.data_ptr(), 64, w1.data_ptr(), 64
    d.dact_aux, d.ld_aux, d.dact = x.data_ptr(), 64, 2
    g = torch.zeros(x.shape[0], 64, device=DEV)
    d.out_f32, d.ldc_f32 = g.data_ptr(), 64
    import ctypes as C
    N.check(N.lib().sunb_gemm(C.byref(d), 0, N.current_stream()), "sunb_gemm")
    torch.cuda.synchronize()
    gref = 0.5 * (1 + torch.erf(xd * 0.70710678118654752)) + xd * torch.exp(-0.5 * xd * xd) * 0.3989422804014327
    assert (g.double() - gref).abs().max().item() < 1e-5


def test_gconv3x3_forward_and_dgrad():
    """Warp-MMA grouped conv (sunb_gconv3x3) vs torch conv2d(groups=8) forward, GELU, pre-activation copy and dgrad."""
    from sunb200 import packing
    B = 3
    x = rnd(B * 400, 256, seed=40).bfloat16()
    w = rnd(256, 32, 3, 3, seed=41, scale=(9 * 32) ** -0.5)
    wg = packing._grouped_taps(w).bfloat16().contiguous()
    y = torch.empty(B * 400, 256, device=DEV, dtype=torch.bfloat16)
    y2 = torch.empty_like(y)
    lib, st = N.lib(), N.current_stream()
    N.check(lib.sunb_gconv3x3(x.data_ptr(), 256, wg.data_ptr(), y.data_ptr(), 256, y2.data_ptr(), 256, None, 0, B, 2, 0, st), "gconv")
    torch.cuda.synchronize()
    xf = x.float().view(B, 20, 20, 256).permute(0, 3, 1, 2).requires_grad_(True)
    pre = F.conv2d(xf, w.bfloat16().float(), padding=1, groups=8)
    ref_pre = pre.permute(0, 2, 3, 1).reshape(-1, 256)
    assert rel_err(y2, ref_pre) < BF16_OUT
    assert rel_err(y, act_ref(ref_pre, 2)) < BF16_OUT
    # dgrad: conv-transpose weights packed on the device, chain-rule factor gelu'(aux)
    dy = rnd(B * 400, 256, seed=42).bfloat16()
    aux = rnd(B * 400, 256, seed=43).bfloat16()
    wd = torch.empty(8 * 9 * 32, 32, device=DEV, dtype=torch.bfloat16)
    N.check(lib.sunb_gconv_pack(w.data_ptr(), wd.data_ptr(), 1, st), "pack")
    wf = torch.empty(8 * 9 * 32, 32, device=DEV, dtype=torch.bfloat16)
    N.check(lib.sunb_gconv_pack(w.data_ptr(), wf.data_ptr(), 0, st), "pack")
    assert torch.equal(wf.view(8, 9, 32, 32), wg)
    dx = torch.empty_like(y)
    N.check(lib.sunb_gconv3x3(dy.data_ptr(), 256, wd.data_ptr(), dx.data_ptr(), 256, None, 0, aux.data_ptr(), 256, B, 0, 2, st), "gconv dgrad")
    torch.cuda.synchronize()
    pre.backward(dy.float().view(B, 20, 20, 256).permute(0, 3, 1, 2))
    a = aux.float()
    gprime = 0.5 * (1 + torch.erf(a * 0.70710678118654752)) + a * torch.exp(-0.5 * a * a) * 0.3989422804014327
    ref_dx = xf.grad.permute(0, 2, 3, 1).reshape(-1, 256) * gprime
    assert rel_err(dx, ref_dx) < 1e-2


@pytest.mark.parametrize("B,s2d", [(1, 0), (3, 1), (301, 0)])
def test_convmlp_tail_fused(B, s2d):
    """sunb_convmlp_tail: out = resid + conv3(gelu(gconv3x3(h1))) in one kernel vs torch (fp32 on the same bf16 operands)."""
    from sunb200 import packing
    h1 = rnd(B * 400, 256, seed=70).bfloat16()
    x = rnd(B * 400, 128, seed=71).bfloat16()
    w2 = rnd(256, 32, 3, 3, seed=72, scale=(9 * 32) ** -0.5)
    w3 = rnd(128, 256, seed=73, scale=256 ** -0.5)
    blob = packing._convmlp_tail_blob(w2, w3).bfloat16().contiguous()
    out = torch.full((B * 400, 128), float("nan"), device=DEV, dtype=torch.bfloat16)
    N.check(N.lib().sunb_convmlp_tail(h1.data_ptr(), blob.data_ptr(), x.data_ptr(), out.data_ptr(), B, s2d, N.current_stream()),
            "sunb_convmlp_tail")
    torch.cuda.synchronize()
    hf = h1.float().view(B, 20, 20, 256).permute(0, 3, 1, 2)
    h2 = act_ref(F.conv2d(hf, w2.bfloat16().float(), padding=1, groups=8), 2).bfloat16().float()     # the kernel rounds h2 to bf16
    y = F.conv2d(h2, w3.bfloat16().float().view(128, 256, 1, 1)).permute(0, 2, 3, 1).reshape(B, 20, 20, 128) + x.float().view(B, 20, 20, 128)
    if s2d:
        y = y.reshape(B, 10, 2, 10, 2, 128).permute(0, 1, 3, 2, 4, 5)
    ref = y.reshape(B * 400, 128)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < BF16_OUT


@pytest.mark.parametrize("B,s2d", [(1, 0), (3, 1), (77, 0), (301, 1)])
def test_mlp_fused(B, s2d):
    """Fused stage-2 MLP (mlp_fused.cu): out = x + conv3(gelu(conv1(x) + b1)) vs the fp32 formula (visformer.py:127-163, 259-263)
    and BIT-EXACT vs the two tcgen05 GEMMs it replaces; ragged last tile (B = 1, 3, 77 -> 100, 300, 7700 rows), in-place
    update and the 2x2 space-to-depth store."""
    M, Cc, H = B * 100, 256, 1024
    x = rnd(M, Cc, seed=31).bfloat16()
    w1 = (rnd(H, Cc, seed=32) * 0.06).bfloat16()
    b1 = rnd(H, seed=33) * 0.1
    w3 = (rnd(Cc, H, seed=34) * 0.03).bfloat16()
    hid = run_gemm(x, w1, M, H, Cc, bias=b1, act=2)
    two = run_gemm(hid, w3, M, Cc, H, resid=x, out_map=s2d, oHW=(10, 10))
    xin = x.clone()
    out = torch.full_like(x, float("nan")) if s2d else xin          # in place unless space-to-depth
    N.check(N.lib().sunb_mlp_fused(xin.data_ptr(), w1.data_ptr(), b1.data_ptr(), w3.data_ptr(), out.data_ptr(), M, s2d, 10, 10,
                                   N.current_stream()), "mlp_fused")
    torch.cuda.synchronize()
    assert torch.equal(out, two)
    ref = x.float() + act_ref(x.float() @ w1.float().t() + b1, 2).bfloat16().float() @ w3.float().t()
    if s2d:
        ref = ref.reshape(B, 5, 2, 5, 2, Cc).permute(0, 1, 3, 2, 4, 5).reshape(M, Cc)
    assert rel_err(out, ref) < BF16_OUT

"""CPU tests of the host-side logic: packing maths, registry/state_dict contract, C-ABI surface."""
import ctypes
import os
import re

import pytest
import torch

import sun_oracle as O
from sunb200 import native as N, packing
import emulate


def test_packing_folds_match_oracle():
    """BN-folded, re-laid-out weights re-applied with plain torch fp32 ops == oracle layer outputs."""
    sd = O.randomize_bn(O.init_meta_baseline_state_dict(12345), seed=7)
    x = O.make_episode_images(101, 2, 1)
    taps_o, taps_e = {}, {}
    with torch.no_grad():
        dense, pooled = O.encoder_forward(sd, x, "encoder.", taps=taps_o)
        P = packing.pack_encoder(sd, "encoder.", wdtype=torch.float32)
        d2, p2 = emulate.emulate_forward(P, x, taps_e)
    for k, ref in taps_o.items():
        ref = ref.permute(0, 2, 3, 1)
        err = (taps_e[k] - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
        assert err < 5e-4, (k, err)      # fp32 re-association only (raw-init activations reach ~550)
    assert (p2 - pooled).abs().max().item() / pooled.abs().max().item() < 5e-4
    assert (d2 - dense.permute(0, 2, 3, 1)).abs().max().item() / dense.abs().max().item() < 5e-4


def test_packed_layouts():
    sd = O.init_encoder_state_dict(3)
    P = packing.pack_encoder(sd)
    assert P["stem_w2"].shape == (9, 128, 64) and P["stem_w2"].dtype == torch.bfloat16
    assert P["s1.0.w2"].shape == (8, 9, 32, 32)
    blk = packing._grouped_pairs(sd["stage1.0.mlp.conv2.weight"])
    assert blk.shape == (4, 9, 64, 64)
    assert (blk[:, :, :32, 32:] == 0).all() and (blk[:, :, 32:, :32] == 0).all()      # block-diagonal pairs
    # heads padded 42 -> 48 and 85 -> 96 channels with zero rows / bias / columns (16-byte aligned head segments)
    assert P["s2.0.wqkv"].shape == (864, 256) and P["s3.0.wqkv"].shape == (1728, 512)
    assert P["s2.0.bqkv"].shape == (864,) and P["s3.0.bqkv"].shape == (1728,)
    for name, d, dp in (("s2.0", 42, 48), ("s3.0", 85, 96)):
        wq = P[name + ".wqkv"].float().reshape(18, dp, -1)
        assert (wq[:, d:] == 0).all() and wq[:, :d].abs().sum() > 0
        assert (P[name + ".bqkv"].reshape(18, dp)[:, d:] == 0).all()
        wp = P[name + ".wproj"].float()
        assert wp.shape[1] == 6 * dp and (wp.reshape(-1, 6, dp)[:, :, d:] == 0).all()
    assert P["pe2_w"].shape == (256, 512) and P["pe2_bias"].shape == (100, 256)
    assert P["pe3_w"].shape == (512, 1024) and P["pe3_bias"].shape == (25, 512)
    for k, v in P.items():
        assert v.is_contiguous(), k


def test_convmlp_tail_blob_layout():
    """Operand blob of the fused stage-1 block tail: no-swizzle K-major core matrices, 26,624 bytes per group."""
    g = torch.Generator().manual_seed(3)
    w2 = torch.randn(256, 32, 3, 3, generator=g)
    w3 = torch.randn(128, 256, generator=g)
    blob = packing._convmlp_tail_blob(w2, w3).reshape(8, -1)
    assert blob.shape[1] * 2 == 26624                      # bytes per group in bf16
    for (grp, tap, n, k) in [(0, 0, 0, 0), (3, 5, 17, 9), (7, 8, 31, 31), (2, 4, 8, 24)]:
        c, j = k // 8, k % 8
        assert blob[grp, ((tap * 4 + c) * 32 + n) * 8 + j] == w2[grp * 32 + n, k, tap // 3, tap % 3]
    for (grp, n, k) in [(0, 0, 0), (5, 100, 13), (7, 127, 31)]:
        c, j = k // 8, k % 8
        assert blob[grp, 9 * 4 * 32 * 8 + (c * 128 + n) * 8 + j] == w3[n, grp * 32 + k]


def test_registry_and_state_dict_contract():
    import models
    m = models.make("meta-baseline", encoder="visformer_micro_80", encoder_args={"drop_path_rate": 0.5})
    ref = O.init_meta_baseline_state_dict(1)
    got = m.state_dict()
    assert list(got.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(got[k].shape) == tuple(ref[k].shape), k
    assert m.encoder.out_dim == 512
    assert sum(p.numel() for p in m.parameters()) == 12_531_393
    m.load_state_dict(ref)
    assert models.make(None) is None
    assert {"meta-baseline", "visformer_micro_80", "visformer", "classifier", "linear-classifier", "nn-classifier",
            "token-label"} <= set(models.models)
    ck = {"model": "meta-baseline", "model_args": {"encoder": "visformer_micro_80", "encoder_args": {}}, "model_sd": ref}
    m2 = models.load(ck)
    assert torch.equal(m2.state_dict()["encoder.stem.conv1.weight"], ref["encoder.stem.conv1.weight"])
    # encoder hand-over used by test_few_shot.py:57-63 and train_meta.py:124-126
    m3 = models.make("meta-baseline", encoder=None)
    m3.encoder = m2.encoder
    m.encoder.load_state_dict(m2.encoder.state_dict())
    import utils
    utils.freeze_bn(m)
    assert all(not b.training for b in m.modules() if isinstance(b, torch.nn.BatchNorm2d))


def test_index_utils_match_oracle():
    import utils.few_shot as fs
    for (way, shot, query, ep) in [(5, 1, 15, 1), (5, 5, 15, 2), (10, 1, 5, 8), (3, 2, 1, 4), (1, 1, 1, 1)]:
        n = ep * way * (shot + query)
        ids = torch.arange(n).view(n, 1, 1, 1).float()
        s, q = fs.split_shot_query(ids, way, shot, query, ep)
        so, qo = O.split_shot_query(ids, way, shot, query, ep)
        assert torch.equal(s, so) and torch.equal(q, qo)
        assert torch.equal(fs.make_nk_label(way, query, ep), O.make_nk_label(way, query, ep))


def test_cpu_tensors_fail_loudly():
    import utils
    with pytest.raises(RuntimeError):
        utils.compute_logits(torch.zeros(1, 2, 4), torch.zeros(1, 3, 4))


def test_c_abi_exports_every_declared_symbol():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "sunb200.h")).read()
    declared = set(re.findall(r"\b(sunb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    assert os.path.exists(N.LIB_PATH), "libsunb200.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert N.lib().sunb_abi_version() == N.ABI_VERSION
    # struct layouts mirror the header (sizes are what the C compiler produces for these field lists)
    assert ctypes.sizeof(N.ConvMlpW) == 40 and ctypes.sizeof(N.AttnBlockW) == 48
    assert ctypes.sizeof(N.EncoderWeights) == 9 * 8 + 4 * 40 + 16 + 2 * 48 + 16 + 3 * 48 + 16


def test_every_pdl_launched_kernel_waits_for_its_predecessor():
    """Launch protocol (csrc/common.cuh): a kernel launched through sunb_launch carries the programmatic-dependent-launch
    attribute, so it MUST execute griddepcontrol.wait (pdl_wait) before its first global access; a kernel without the wait
    would race with the previous kernel of the stream.  Source-level audit of csrc/*.cu."""
    import glob
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "few-shot-vit_b200", "csrc")
    sources = {f: open(f).read() for f in glob.glob(os.path.join(csrc, "*.cu"))}
    launched = {(f, m.group(1)) for f, s in sources.items() for m in re.finditer(r"sunb_launch\(&([A-Za-z_0-9]+)", s)}
    assert len(launched) >= 30
    for f, name in sorted(launched):
        s = sources[f]
        m = re.search(r"__global__[^;{]*?\b" + name + r"\s*\(", s, re.S)
        assert m, f"{name}: kernel definition not found in {os.path.basename(f)}"
        i, depth = m.end() - 1, 0
        while True:                                   # end of the parameter list
            depth += s[i] == "("
            depth -= s[i] == ")"
            if depth == 0:
                break
            i += 1
        body_start = s.index("{", i)
        nxt = s.find("__global__", body_start)
        body = s[body_start: nxt if nxt > 0 else len(s)]
        w = body.find("pdl_wait()")
        assert w >= 0, f"{name} is launched with the PDL attribute but never calls pdl_wait()"
        pre = body[:w]
        assert "__ldg" not in pre and "ld_global" not in pre and "tma_load" not in pre, f"{name}: global access before pdl_wait()"


def _emulate_pack(entry):
    """Reference semantics of one SunbPackDesc (include/sunb200.h): gather with strides, zero padding."""
    key, src, dims, strides, off, ldd, cols, valid2 = entry
    flat = src.detach().reshape(-1)
    a, b, c, d = (torch.arange(n) for n in dims)
    idx = (off + a[:, None, None, None] * strides[0] + b[None, :, None, None] * strides[1] + c[None, None, :, None] * strides[2]
           + d[None, None, None, :] * strides[3])
    out = torch.zeros(dims[0], dims[1], dims[2], ldd)
    c_ok = dims[2] if not valid2 else valid2                          # the kernel never reads the source for c >= valid2
    assert int(idx[:, :, :c_ok].min()) >= 0 and int(idx[:, :, :c_ok].max()) < flat.numel(), f"{key}: plan reads out of bounds"
    out[:, :, :c_ok, : dims[3]] = flat[idx[:, :, :c_ok]]
    return out.reshape(-1, cols)


def test_weight_pack_plan_layouts():
    """The table that drives sunb_pack_weights (train.py::weight_pack_plan) produces, for every GEMM / conv weight, the forward
    operand [N][K] and the data-gradient operand [K][N] the kernels expect: taps mirrored for 3x3 dgrads, grouped weights as
    [group][tap][n][k], heads padded from d to 48 / 96 channels with zero rows / columns (reference shapes: visformer.py:146-194)."""
    from sunb200 import train as T
    g = torch.Generator().manual_seed(5)
    sd = O.init_meta_baseline_state_dict(12345)
    P = {k[len("encoder."):]: torch.randn(v.shape, generator=g) for k, v in sd.items()
         if k.startswith("encoder.") and v.dtype == torch.float32 and v.dim() >= 2}
    plan = {e[0]: e for e in T.weight_pack_plan(P)}
    assert len(plan) == 72
    got = {k: _emulate_pack(e) for k, e in plan.items()}
    # dense 3x3: forward [tap][n][c]; dgrad = transposed conv: [tap'][c][n] with tap' = 8 - tap
    w = P["stem.conv3.weight"]                                        # [n, c, 3, 3]
    assert torch.equal(got["stem.conv3.f"], w.permute(2, 3, 0, 1).reshape(9 * 128, 128))
    assert torch.equal(got["stem.conv3.d"], w.flip(2, 3).permute(2, 3, 1, 0).reshape(9 * 128, 128))
    w = P["stem.conv2.weight"]
    assert torch.equal(got["stem.conv2.d"], w.flip(2, 3).permute(2, 3, 1, 0).reshape(9 * 64, 128))
    # 1x1 convs: forward = the weight matrix, dgrad = its transpose
    w = P["stage1.0.mlp.conv1.weight"].reshape(256, 128)
    assert torch.equal(got["stage1.0.mlp.conv1.f"], w) and torch.equal(got["stage1.0.mlp.conv1.d"], w.t())
    w = P["stage3.2.mlp.conv3.weight"].reshape(512, 2048)
    assert torch.equal(got["stage3.2.conv3.f"], w) and torch.equal(got["stage3.2.conv3.d"], w.t())
    # grouped 3x3 [256, 32, 3, 3] -> [group][tap][n][k]; dgrad swaps n / k and mirrors the taps
    w = P["stage1.1.mlp.conv2.weight"].reshape(8, 32, 32, 9)          # [group, n, k, tap]
    assert torch.equal(got["stage1.1.mlp.conv2.f"], w.permute(0, 3, 1, 2).reshape(-1, 32))
    assert torch.equal(got["stage1.1.mlp.conv2.d"], w.flip(3).permute(0, 3, 2, 1).reshape(-1, 32))
    # PatchEmbed 2x2 stride 2: forward [n][(tap, c)], dgrad [(tap, c)][n]
    w = P["patch_embed2.proj.weight"]                                 # [256, 128, 2, 2]
    assert torch.equal(got["pe2.f"], w.permute(0, 2, 3, 1).reshape(256, 512))
    assert torch.equal(got["pe2.d"], w.permute(2, 3, 1, 0).reshape(512, 256))
    # attention: heads padded d -> ds with zeros
    for stage, dim, d, ds in (("2", 256, 42, 48), ("3", 512, 85, 96)):
        wq = P[f"stage{stage}.0.attn.qkv.weight"].reshape(18, d, dim)
        pad = torch.zeros(18, ds, dim)
        pad[:, :d] = wq
        assert torch.equal(got[f"stage{stage}.0.qkv.f"], pad.reshape(18 * ds, dim))
        assert torch.equal(got[f"stage{stage}.0.qkv.d"], pad.reshape(18 * ds, dim).t())
        wp = P[f"stage{stage}.0.attn.proj.weight"].reshape(dim, 6, d)
        padp = torch.zeros(dim, 6, ds)
        padp[:, :, :d] = wp
        assert torch.equal(got[f"stage{stage}.0.proj.f"], padp.reshape(dim, 6 * ds))
        assert torch.equal(got[f"stage{stage}.0.proj.d"], padp.reshape(dim, 6 * ds).t())

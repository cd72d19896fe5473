"""Weight packing for the eval-mode encoder: BatchNorm folding, layout permutation, bf16 conversion.

Pure torch tensor algebra (device agnostic), so the folding / layout logic is unit-tested on the CPU against
the oracle (tests/emulate.py re-applies the packed weights with plain torch ops) before any kernel runs.  Layouts match include/sunb200.h (SunbEncoderWeights).

Folding rules:
  norm -> 1x1 conv (blocks):  conv(BN(x)) = (W diag(s)) x + W t         s = gamma/sqrt(var+eps), t = beta - mean*s
  conv -> BN (stem, PatchEmbed): BN(conv(x)+b) = diag(s) W x + (s (b - mean) + beta)
Reference: test_phase/models/visformer.py:118-124 (BatchNorm), :254-262 (Block), :209-216 (stem), :276-287 (PatchEmbed).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

HEADS = 6
BN_EPS = 1e-5
DEPTH = (4, 2, 3)


def bn_affine(sd: Dict[str, torch.Tensor], p: str):
    """Eval-mode BatchNorm as y = x*s + t."""
    s = sd[p + ".weight"].float() / torch.sqrt(sd[p + ".running_var"].float() + BN_EPS)
    t = sd[p + ".bias"].float() - sd[p + ".running_mean"].float() * s
    return s, t


def _conv3x3_taps(w: torch.Tensor) -> torch.Tensor:
    """[N, C, 3, 3] -> [9, N, C] with tap = ky*3 + kx."""
    return w.permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1]).contiguous()


def _grouped_pairs(w: torch.Tensor, groups: int = 8) -> torch.Tensor:
    """Grouped 3x3 weight [256, 32, 3, 3] (8 groups) -> [4 pairs][9 taps][64 n][64 k] block-diagonal."""
    n_out, cpg = w.shape[0], w.shape[1]
    opg = n_out // groups
    pairs = groups // 2
    out = torch.zeros(pairs, 9, 2 * opg, 2 * cpg, dtype=w.dtype, device=w.device)
    taps = w.permute(2, 3, 0, 1).reshape(9, n_out, cpg)
    for g in range(groups):
        p, h = g // 2, g % 2
        out[p, :, h * opg:(h + 1) * opg, h * cpg:(h + 1) * cpg] = taps[:, g * opg:(g + 1) * opg, :]
    return out.contiguous()


def _grouped_taps(w: torch.Tensor, groups: int = 8) -> torch.Tensor:
    """Grouped 3x3 weight [256, 32, 3, 3] -> [8 groups][9 taps][32 n][32 k] (operand of sunb_gconv3x3)."""
    n_out, cpg = w.shape[0], w.shape[1]
    opg = n_out // groups
    return w.reshape(groups, opg, cpg, 9).permute(0, 3, 1, 2).contiguous()


def _convmlp_tail_blob(w2: torch.Tensor, w3: torch.Tensor, groups: int = 8) -> torch.Tensor:
    """Operand blob of sunb_convmlp_tail: per group g [grouped taps: 9 x 4 k-chunks x 32 n x 8 k | conv3 slice: 4 k-chunks x
    128 n x 8 k] -- both no-swizzle K-major UMMA operands (core matrix = 8 rows x 16 bytes), 26,624 bytes per group.
    w2 [256, 32, 3, 3] grouped 3x3 weight, w3 [128, 256] (conv3 as a matrix); returns a flat tensor of w2's dtype."""
    n_out, cpg = w2.shape[0], w2.shape[1]
    opg = n_out // groups
    taps = w2.reshape(groups, opg, cpg, 9).permute(0, 3, 1, 2)                      # [g][tap][n][k]
    a = taps.reshape(groups, 9, opg, cpg // 8, 8).permute(0, 1, 3, 2, 4)            # [g][tap][c][n][8]
    b = w3.reshape(w3.shape[0], groups, cpg // 8, 8).permute(1, 2, 0, 3)            # [g][c][n][8]
    return torch.cat([a.reshape(groups, -1), b.reshape(groups, -1)], dim=1).contiguous().reshape(-1)


def _pad_cols(w: torch.Tensor, mult: int = 8) -> torch.Tensor:
    k = w.shape[1]
    kp = (k + mult - 1) // mult * mult
    return w if kp == k else F.pad(w, (0, kp - k))


HEAD_PAD = {42: 48, 85: 96}      # head_dim -> padded head stride of the eval engine (16-byte aligned head segments)


def _pad_heads_rows(w: torch.Tensor, d: int, dp: int, blocks: int) -> torch.Tensor:
    """[blocks*d, ...] -> [blocks*dp, ...]: every block of d rows is followed by dp - d zero rows."""
    w = w.reshape(blocks, d, *w.shape[1:])
    out = w.new_zeros(blocks, dp, *w.shape[2:])
    out[:, :d] = w
    return out.reshape(blocks * dp, *w.shape[2:])


def pack_encoder(sd: Dict[str, torch.Tensor], prefix: str = "", wdtype: torch.dtype = torch.bfloat16) -> Dict[str, torch.Tensor]:
    """state_dict (reference names) -> dict of packed tensors keyed like SunbEncoderWeights fields.
    `wdtype=torch.float32` keeps GEMM weights in fp32 (CPU emulation tests)."""
    g = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    P: Dict[str, torch.Tensor] = {}

    def w2d(name):
        return g[name].float().flatten(1)

    # stem: conv -> BN folds (scale output rows)
    s, t = bn_affine(g, "stem.bn1")
    P["stem_w1"] = (w2d("stem.conv1.weight") * s[:, None]).contiguous()
    P["stem_b1"] = t.contiguous()
    s, t = bn_affine(g, "stem.downsample.1")
    P["stem_wd"] = (w2d("stem.downsample.0.weight") * s[:, None]).contiguous()
    P["stem_bd"] = t.contiguous()
    s, t = bn_affine(g, "stem.bn2")
    P["stem_w2"] = _conv3x3_taps(g["stem.conv2.weight"].float() * s[:, None, None, None]).to(wdtype)
    P["stem_b2"] = t.contiguous()
    s, t = bn_affine(g, "stem.bn3")
    P["stem_w3"] = _conv3x3_taps(g["stem.conv3.weight"].float() * s[:, None, None, None]).to(wdtype)
    P["stem_b3"] = t.contiguous()
    P["pos1"] = g["pos_embed1"][0].float().permute(1, 2, 0).reshape(400, 128).contiguous()

    for i in range(DEPTH[0]):
        b = f"stage1.{i}."
        s, t = bn_affine(g, b + "norm2.bn")
        w1 = w2d(b + "mlp.conv1.weight")
        P[f"s1.{i}.w1"] = (w1 * s[None, :]).to(wdtype).contiguous()
        P[f"s1.{i}.b1"] = (w1 @ t).contiguous()
        P[f"s1.{i}.w2"] = _grouped_taps(g[b + "mlp.conv2.weight"].float()).to(wdtype)
        P[f"s1.{i}.w3"] = w2d(b + "mlp.conv3.weight").to(wdtype).contiguous()
        P[f"s1.{i}.w23"] = _convmlp_tail_blob(g[b + "mlp.conv2.weight"].float(), w2d(b + "mlp.conv3.weight")).to(wdtype)

    for stage, depth, hw in (("2", DEPTH[1], 100), ("3", DEPTH[2], 25)):
        pe = f"patch_embed{stage}."
        s, t = bn_affine(g, pe + "norm.bn")
        w = g[pe + "proj.weight"].float() * s[:, None, None, None]            # [N, C, 2, 2]
        P[f"pe{stage}_w"] = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(wdtype).contiguous()   # k = (dy, dx, c)
        bias = g[pe + "proj.bias"].float() * s + t                             # s*(b - mean) + beta
        pos = g[f"pos_embed{stage}"][0].float()                                # [N, h, w]
        P[f"pe{stage}_bias"] = (bias[None, :] + pos.permute(1, 2, 0).reshape(hw, -1)).contiguous()
        for i in range(depth):
            b = f"stage{stage}.{i}."
            s, t = bn_affine(g, b + "norm1.bn")
            wq = w2d(b + "attn.qkv.weight")                       # [3*6*d, C]
            d = wq.shape[0] // (3 * HEADS)
            dp = HEAD_PAD[d]
            P[f"s{stage}.{i}.wqkv"] = _pad_heads_rows(wq * s[None, :], d, dp, 3 * HEADS).to(wdtype).contiguous()
            P[f"s{stage}.{i}.bqkv"] = _pad_heads_rows(wq @ t, d, dp, 3 * HEADS).contiguous()
            wp = w2d(b + "attn.proj.weight")                      # [C, 6*d] -> zero columns at the head pads
            P[f"s{stage}.{i}.wproj"] = _pad_heads_rows(wp.t().contiguous(), d, dp, HEADS).t().to(wdtype).contiguous()
            s, t = bn_affine(g, b + "norm2.bn")
            w1 = w2d(b + "mlp.conv1.weight")
            P[f"s{stage}.{i}.w1"] = (w1 * s[None, :]).to(wdtype).contiguous()
            P[f"s{stage}.{i}.b1"] = (w1 @ t).contiguous()
            P[f"s{stage}.{i}.w3"] = w2d(b + "mlp.conv3.weight").to(wdtype).contiguous()
    s, t = bn_affine(g, "norm.bn")
    P["final_scale"], P["final_shift"] = s.contiguous(), t.contiguous()
    return P


def to_struct(P: Dict[str, torch.Tensor]):
    """Packed dict (CUDA tensors) -> ctypes SunbEncoderWeights.  The dict must outlive the struct."""
    from . import native as N
    w = N.EncoderWeights()
    for f in ("stem_w1", "stem_b1", "stem_wd", "stem_bd", "stem_w2", "stem_b2", "stem_w3", "stem_b3", "pos1",
              "pe2_w", "pe2_bias", "pe3_w", "pe3_bias", "final_scale", "final_shift"):
        setattr(w, f, P[f].data_ptr())
    for i in range(4):
        for f in ("w1", "b1", "w2", "w3", "w23"):
            setattr(w.s1[i], f, P[f"s1.{i}.{f}"].data_ptr())
    for stage, arr, depth in (("2", w.s2, 2), ("3", w.s3, 3)):
        for i in range(depth):
            for f in ("wqkv", "bqkv", "wproj", "w1", "b1", "w3"):
                setattr(arr[i], f, P[f"s{stage}.{i}.{f}"].data_ptr())
    return w

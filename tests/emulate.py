"""Torch emulation of the packed eval plan (same operand layouts and row mappings as csrc/api.cu).
TEST INFRASTRUCTURE: validates the BN folding / layout permutations of sunb200.packing on the CPU."""
from typing import Dict

import torch
import torch.nn.functional as F

from sunb200.packing import HEADS, HEAD_PAD

def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * 0.70710678118654752))


def _conv3x3_nhwc(x: torch.Tensor, taps: torch.Tensor) -> torch.Tensor:
    """x [B,H,W,C], taps [9,N,C] -> [B,H,W,N]: nine shifted GEMMs with zero padding (the kernel's K loop)."""
    B, H, W, C = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    out = torch.zeros(B, H, W, taps.shape[1], dtype=x.dtype)
    for tap in range(9):
        dy, dx = tap // 3, tap % 3
        out += xp[:, dy:dy + H, dx:dx + W, :] @ taps[tap].t()
    return out


def _s2d(x: torch.Tensor) -> torch.Tensor:
    """[B,H,W,C] raster -> [B*(H/2)*(W/2), 4*C] rows in the kernel's MAP_S2D order, k = (dy, dx, c)."""
    B, H, W, C = x.shape
    return x.reshape(B, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B * (H // 2) * (W // 2), 4 * C)


def emulate_forward(P: Dict[str, torch.Tensor], x: torch.Tensor, taps: Dict[str, torch.Tensor] = None):
    """fp32 emulation of sunb_encoder_forward from packed weights.  Returns (dense NHWC [B,5,5,512], pooled)."""
    f = {k: v.float() for k, v in P.items()}
    B = x.shape[0]
    a1 = F.conv2d(x, f["stem_w1"].reshape(64, 3, 3, 3), f["stem_b1"], stride=2, padding=1)
    a1 = F.leaky_relu(a1, 0.1).permute(0, 2, 3, 1)
    idn = F.conv2d(x, f["stem_wd"].reshape(128, 3, 3, 3), f["stem_bd"], stride=2, padding=1).permute(0, 2, 3, 1)
    a2 = F.leaky_relu(_conv3x3_nhwc(a1, f["stem_w2"]) + f["stem_b2"], 0.1)
    c3 = F.leaky_relu(_conv3x3_nhwc(a2, f["stem_w3"]) + f["stem_b3"] + idn, 0.1)
    s1 = c3.reshape(B, 20, 2, 20, 2, 128).amax(dim=(2, 4)) + f["pos1"].reshape(1, 20, 20, 128)

    def tap(name, t):
        if taps is not None:
            taps[name] = t.clone()
    tap("stem", s1)
    for i in range(4):
        h1 = _gelu(s1 @ f[f"s1.{i}.w1"].t() + f[f"s1.{i}.b1"])
        h2 = torch.empty_like(h1)
        for gi in range(8):
            h2[..., gi * 32:(gi + 1) * 32] = _gelu(_conv3x3_nhwc(h1[..., gi * 32:(gi + 1) * 32], f[f"s1.{i}.w2"][gi]))
        s1 = s1 + h2 @ f[f"s1.{i}.w3"].t()
        tap(f"stage1.{i}", s1)
    t = _s2d(s1) @ f["pe2_w"].t()
    t = (t.reshape(B, 100, 256) + f["pe2_bias"]).reshape(B, 10, 10, 256)
    tap("patch_embed2", t)

    def attn_block(t, pre, d):
        Bb, H, W, Cc = t.shape
        S = H * W
        x2 = t.reshape(Bb * S, Cc)
        dp = HEAD_PAD[d]
        qkv = x2 @ f[pre + "wqkv"].t() + f[pre + "bqkv"]                 # padded heads: pad channels are exact zeros
        qkv = qkv.reshape(Bb, S, 3, HEADS, dp).permute(2, 0, 3, 1, 4)
        pr = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) * d ** -0.5, dim=-1)
        ao = (pr @ qkv[2]).permute(0, 2, 1, 3).reshape(Bb * S, HEADS * dp)
        x2 = x2 + ao @ f[pre + "wproj"].t()
        hid = _gelu(x2 @ f[pre + "w1"].t() + f[pre + "b1"])
        return (x2 + hid @ f[pre + "w3"].t()).reshape(Bb, H, W, Cc)

    for i in range(2):
        t = attn_block(t, f"s2.{i}.", 42)
        tap(f"stage2.{i}", t)
    t = _s2d(t) @ f["pe3_w"].t()
    t = (t.reshape(B, 25, 512) + f["pe3_bias"]).reshape(B, 5, 5, 512)
    tap("patch_embed3", t)
    for i in range(3):
        t = attn_block(t, f"s3.{i}.", 85)
        tap(f"stage3.{i}", t)
    dense = t * f["final_scale"] + f["final_shift"]
    return dense, dense.mean(dim=(1, 2))

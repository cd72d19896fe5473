"""CPU oracle for the SUN episodic Visformer hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (functional, state_dict-driven, torch fp32 on the CPU) of the
reference algorithm.  It is never imported by the product package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it,
and there only as the checker / the timed CPU arm.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference modules from
``/root/reference`` (in the authoring container), loads the state_dict produced by
``init_state_dict`` below into them, and stores the reference's own outputs under
``tests/golden/``.  ``tests/test_oracle_golden.py`` checks every function here against those files.
(The reference itself ships no golden vectors or known-answer tests)

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------------------
# Architecture constants of 'visformer_micro_80' (test_phase/models/visformer.py:482-487)
# --------------------------------------------------------------------------------------
IMG = 80
STEM_CH = 64          # init_channels
DIM1, DIM2, DIM3 = 128, 256, 512   # embed_dim//2, embed_dim, embed_dim*2
DEPTH = (4, 2, 3)
HEADS = 6
GROUPS = 8
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LRELU_SLOPE = 0.1
OUT_DIM = 512


def head_dim(dim: int, ratio: float = 1.0) -> int:
    """visformer.py:172  head_dim = round(dim // num_heads * ratio) -> 42 (dim 256) / 85 (dim 512)."""
    return round(dim // HEADS * ratio)


def drop_path_rates(drop_path_rate: float) -> Sequence[float]:
    """visformer.py:312  one rate per block, linspace(0, rate, 9)."""
    return [x.item() for x in torch.linspace(0, drop_path_rate, sum(DEPTH))]


# --------------------------------------------------------------------------------------
# state_dict: names / shapes (probed from the reference) and a seeded init
# --------------------------------------------------------------------------------------
def _bn_entries(prefix: str, c: int):
    return [(prefix + ".weight", (c,)), (prefix + ".bias", (c,)),
            (prefix + ".running_mean", (c,)), (prefix + ".running_var", (c,)),
            (prefix + ".num_batches_tracked", ())]


def encoder_entries():
    """(name, shape) of the 148 encoder state_dict entries, in the reference module order."""
    e = [("pos_embed1", (1, DIM1, 20, 20)), ("pos_embed2", (1, DIM2, 10, 10)), ("pos_embed3", (1, DIM3, 5, 5))]
    e += [("stem.conv1.weight", (STEM_CH, 3, 3, 3))] + _bn_entries("stem.bn1", STEM_CH)
    e += [("stem.conv2.weight", (DIM1, STEM_CH, 3, 3))] + _bn_entries("stem.bn2", DIM1)
    e += [("stem.conv3.weight", (DIM1, DIM1, 3, 3))] + _bn_entries("stem.bn3", DIM1)
    e += [("stem.downsample.0.weight", (DIM1, 3, 3, 3))] + _bn_entries("stem.downsample.1", DIM1)
    for i in range(DEPTH[0]):
        p = f"stage1.{i}."
        hid = DIM1 * 2
        e += _bn_entries(p + "norm2.bn", DIM1)
        e += [(p + "mlp.conv1.weight", (hid, DIM1, 1, 1)),
              (p + "mlp.conv2.weight", (hid, hid // GROUPS, 3, 3)),
              (p + "mlp.conv3.weight", (DIM1, hid, 1, 1))]
    for stage, dim, cin, depth in (("2", DIM2, DIM1, DEPTH[1]), ("3", DIM3, DIM2, DEPTH[2])):
        e += [(f"patch_embed{stage}.proj.weight", (dim, cin, 2, 2)), (f"patch_embed{stage}.proj.bias", (dim,))]
        e += _bn_entries(f"patch_embed{stage}.norm.bn", dim)
        hd = head_dim(dim)
        for i in range(depth):
            p = f"stage{stage}.{i}."
            e += _bn_entries(p + "norm1.bn", dim)
            e += [(p + "attn.qkv.weight", (3 * HEADS * hd, dim, 1, 1)),
                  (p + "attn.proj.weight", (dim, HEADS * hd, 1, 1))]
            e += _bn_entries(p + "norm2.bn", dim)
            e += [(p + "mlp.conv1.weight", (4 * dim, dim, 1, 1)),
                  (p + "mlp.conv3.weight", (dim, 4 * dim, 1, 1))]
    e += _bn_entries("norm.bn", DIM3)
    return e


def _trunc_normal(shape, std: float, gen: torch.Generator) -> Tensor:
    """Same distribution as visformer.py:13-45 (inverse-CDF truncated normal, cut at +-2 absolute)."""
    lo = (1.0 + math.erf((-2.0 / std) / math.sqrt(2.0))) / 2.0
    hi = (1.0 + math.erf((2.0 / std) / math.sqrt(2.0))) / 2.0
    u = torch.rand(shape, generator=gen, dtype=torch.float64)
    t = (2 * lo - 1) + u * ((2 * hi - 1) - (2 * lo - 1))
    t = torch.erfinv(t) * (std * math.sqrt(2.0))
    return t.clamp_(-2.0, 2.0).float()


def init_encoder_state_dict(seed: int = 12345, prefix: str = "") -> SD:
    """Random-init weights with the reference's *distributions* (visformer.py:398-422: conv_init=True
    => kaiming_normal_(fan_out, relu) for every conv, conv bias 0, BN gamma 1 / beta 0, running stats
    0 / 1, pos-embeds trunc-normal(0.02)).  The bit stream is this oracle's own (one Generator, tensors
    drawn in `encoder_entries` order) so that it is reproducible on any box without the reference; the
    golden files were produced by loading exactly these tensors into the reference modules."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for name, shape in encoder_entries():
        leaf = name.split(".")[-1]
        if name.startswith("pos_embed"):
            t = _trunc_normal(shape, 0.02, gen)
        elif leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.long)
        elif leaf == "running_var":
            t = torch.ones(shape)
        elif leaf in ("running_mean",):
            t = torch.zeros(shape)
        elif len(shape) == 4:                       # conv weight: std = sqrt(2 / (out_ch * kh * kw))
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=gen) * math.sqrt(2.0 / fan_out)
        elif leaf == "weight":                      # BN gamma
            t = torch.ones(shape)
        else:                                       # BN beta / conv bias
            t = torch.zeros(shape)
        sd[prefix + name] = t
    return sd


def init_meta_baseline_state_dict(seed: int = 12345, temp: float = 10.0) -> SD:
    """'meta-baseline' state_dict: `temp` + `encoder.*` (meta_baseline.py:13-22)."""
    sd = {"temp": torch.tensor(temp)}
    sd.update(init_encoder_state_dict(seed, prefix="encoder."))
    return sd


def init_token_label_state_dict(seed: int = 12345, n_classes: int = 64) -> SD:
    """'token-label' state_dict: encoder + classifier Linear(512,n) + classifier_local Linear(512,n+1)
    (sun_meta_training/models/token_label.py:36-46).  nn.Linear default init distribution."""
    sd = init_encoder_state_dict(seed, prefix="encoder.")
    gen = torch.Generator().manual_seed(seed + 1)
    bound = 1.0 / math.sqrt(OUT_DIM)
    for name, n in (("classifier", n_classes), ("classifier_local", n_classes + 1)):
        sd[f"{name}.linear.weight"] = (torch.rand((n, OUT_DIM), generator=gen) * 2 - 1) * bound
        sd[f"{name}.linear.bias"] = (torch.rand((n,), generator=gen) * 2 - 1) * bound
    return sd


def randomize_bn(sd: SD, seed: int = 7, prefix: str = "") -> SD:
    """Give every BatchNorm non-trivial affine parameters and running statistics (a fixture that makes
    BN-folding mistakes visible; raw init has gamma=1, beta=0, mean=0, var=1 which hides them)."""
    gen = torch.Generator().manual_seed(seed)
    out = dict(sd)
    for k in list(sd.keys()):
        if not k.startswith(prefix) or not k.endswith("running_var"):
            continue
        base = k[: -len("running_var")]
        c = sd[k].numel()
        out[base + "running_var"] = 0.5 + torch.rand(c, generator=gen)
        out[base + "running_mean"] = 0.2 * torch.randn(c, generator=gen)
        out[base + "weight"] = 0.75 + 0.5 * torch.rand(c, generator=gen)
        out[base + "bias"] = 0.1 * torch.randn(c, generator=gen)
    for k in list(sd.keys()):
        if k.endswith("proj.bias"):
            out[k] = 0.1 * torch.randn(sd[k].shape, generator=gen)
    return out


# --------------------------------------------------------------------------------------
# Layers
# --------------------------------------------------------------------------------------
class BNState:
    """Collects the running-stat updates of a train-mode forward (torch BatchNorm2d semantics)."""

    def __init__(self):
        self.updates: SD = {}


def batch_norm(sd: SD, p: str, x: Tensor, training: bool, bn_state: Optional[BNState] = None,
               momentum: Optional[float] = BN_MOMENTUM) -> Tensor:
    """nn.BatchNorm2d(eps=1e-5, momentum=0.1) (visformer.py:118-124, 200-201).
    eval: running stats.  train: batch mean / biased var for normalisation; running stats updated with
    the unbiased var (recorded in `bn_state.updates`, the input dict is not mutated)."""
    w, b = sd[p + ".weight"], sd[p + ".bias"]
    rm, rv = sd[p + ".running_mean"], sd[p + ".running_var"]
    if not training:
        scale = w / torch.sqrt(rv + BN_EPS)
        shift = b - rm * scale
        return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    if bn_state is not None:
        n = x.numel() // x.shape[1]
        nbt = int(sd[p + ".num_batches_tracked"]) + 1
        m = momentum if momentum is not None else 1.0 / nbt
        bn_state.updates[p + ".running_mean"] = (1 - m) * rm + m * mean.detach()
        bn_state.updates[p + ".running_var"] = (1 - m) * rv + m * (var.detach() * n / max(n - 1, 1))
        bn_state.updates[p + ".num_batches_tracked"] = torch.tensor(nbt)
    xn = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
    return xn * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def gelu(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (visformer.py:129,145,149)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def apply_drop_path(y: Tensor, mask: Optional[Tensor], keep: float) -> Tensor:
    """visformer.py:89-97 with the Bernoulli draw injected: y / keep * mask[B,1,1,1] (mask in {0,1})."""
    if mask is None:
        return y
    return y / keep * mask.view(-1, 1, 1, 1).to(y.dtype)


def stem(sd: SD, p: str, x: Tensor, training: bool, bn_state=None, momentum=BN_MOMENTUM) -> Tensor:
    """ConvBlock (visformer.py:202-239): three 3x3 convs (first stride 2) with BN + LeakyReLU(0.1),
    a strided 3x3 shortcut with BN, add, LeakyReLU, 2x2 max-pool."""
    o = F.conv2d(x, sd[p + "conv1.weight"], stride=2, padding=1)
    o = F.leaky_relu(batch_norm(sd, p + "bn1", o, training, bn_state, momentum), LRELU_SLOPE)
    o = F.conv2d(o, sd[p + "conv2.weight"], padding=1)
    o = F.leaky_relu(batch_norm(sd, p + "bn2", o, training, bn_state, momentum), LRELU_SLOPE)
    o = F.conv2d(o, sd[p + "conv3.weight"], padding=1)
    o = batch_norm(sd, p + "bn3", o, training, bn_state, momentum)
    idn = F.conv2d(x, sd[p + "downsample.0.weight"], stride=2, padding=1)
    idn = batch_norm(sd, p + "downsample.1", idn, training, bn_state, momentum)
    o = F.leaky_relu(o + idn, LRELU_SLOPE)
    return F.max_pool2d(o, 2)


def conv_mlp(sd: SD, p: str, x: Tensor, spatial: bool) -> Tensor:
    """Mlp (visformer.py:127-163): 1x1 -> GELU -> [grouped 3x3 (8 groups) -> GELU] -> 1x1; dropouts are p=0."""
    h = gelu(F.conv2d(x, sd[p + "conv1.weight"]))
    if spatial:
        h = gelu(F.conv2d(h, sd[p + "conv2.weight"], padding=1, groups=GROUPS))
    return F.conv2d(h, sd[p + "conv3.weight"])


def attention(sd: SD, p: str, x: Tensor, return_probs: bool = False):
    """Attention (visformer.py:166-194): bias-free 1x1 qkv conv, channel order (qkv, head, d),
    tokens row-major (h w), softmax(q k^T * d^-0.5) v, channels (head d), 1x1 proj."""
    B, C, H, W = x.shape
    d = head_dim(C)
    qkv = F.conv2d(x, sd[p + "qkv.weight"])                       # [B, 3*6*d, H, W]
    qkv = qkv.view(B, 3, HEADS, d, H * W).permute(1, 0, 2, 4, 3)    # [3, B, heads, HW, d]
    q, k, v = qkv[0], qkv[1], qkv[2]
    probs = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) * (d ** -0.5), dim=-1)
    o = torch.matmul(probs, v)                                     # [B, heads, HW, d]
    o = o.permute(0, 1, 3, 2).reshape(B, HEADS * d, H, W)
    o = F.conv2d(o, sd[p + "proj.weight"])
    return (o, probs) if return_probs else o


def block(sd: SD, p: str, x: Tensor, attn: bool, spatial: bool, training: bool, bn_state=None,
          momentum=BN_MOMENTUM, dp_masks: Optional[Sequence[Optional[Tensor]]] = None, keep: float = 1.0) -> Tensor:
    """Block.forward (visformer.py:259-263): x + dp(attn(norm1 x)); x + dp(mlp(norm2 x)).
    `dp_masks` = the per-sample DropPath draws for this block in call order (attn first, then mlp)."""
    masks = list(dp_masks) if dp_masks is not None else []
    if attn:
        y = attention(sd, p + "attn.", batch_norm(sd, p + "norm1.bn", x, training, bn_state, momentum))
        x = x + apply_drop_path(y, masks.pop(0) if masks else None, keep)
    y = conv_mlp(sd, p + "mlp.", batch_norm(sd, p + "norm2.bn", x, training, bn_state, momentum), spatial)
    return x + apply_drop_path(y, masks.pop(0) if masks else None, keep)


def patch_embed(sd: SD, p: str, x: Tensor, training: bool, bn_state=None, momentum=BN_MOMENTUM) -> Tensor:
    """PatchEmbed (visformer.py:266-288): 2x2 stride-2 conv with bias, then BN."""
    y = F.conv2d(x, sd[p + "proj.weight"], sd[p + "proj.bias"], stride=2)
    return batch_norm(sd, p + "norm.bn", y, training, bn_state, momentum)


def encoder_forward(sd: SD, x: Tensor, prefix: str = "", training: bool = False,
                    bn_state: Optional[BNState] = None, momentum: Optional[float] = BN_MOMENTUM,
                    drop_path_rate: float = 0.0, dp_masks: Optional[Dict[str, Sequence[Tensor]]] = None,
                    taps: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """Visformer.forward (visformer.py:424-462) for 'visformer_micro_80'.
    Returns (dense[B,512,5,5] after the final BN, pooled[B,512]); test_phase returns only `pooled`
    (:462), sun_meta_training returns both (sun_meta_training/models/visformer.py:464).
    `dp_masks[block_name]` optionally injects DropPath draws (train mode); `taps` collects the layer
    boundaries used by the per-layer GPU tests."""
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    rates = drop_path_rates(drop_path_rate)
    dp_masks = dp_masks or {}

    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach().clone()

    x = stem(sub, "stem.", x, training, bn_state, momentum)
    x = x + sub["pos_embed1"]
    tap("stem", x)
    bi = 0
    for i in range(DEPTH[0]):
        name = f"stage1.{i}"
        x = block(sub, name + ".", x, False, True, training, bn_state, momentum, dp_masks.get(name), 1.0 - rates[bi])
        tap(name, x)
        bi += 1
    x = patch_embed(sub, "patch_embed2.", x, training, bn_state, momentum) + sub["pos_embed2"]
    tap("patch_embed2", x)
    for i in range(DEPTH[1]):
        name = f"stage2.{i}"
        x = block(sub, name + ".", x, True, False, training, bn_state, momentum, dp_masks.get(name), 1.0 - rates[bi])
        tap(name, x)
        bi += 1
    x = patch_embed(sub, "patch_embed3.", x, training, bn_state, momentum) + sub["pos_embed3"]
    tap("patch_embed3", x)
    for i in range(DEPTH[2]):
        name = f"stage3.{i}"
        x = block(sub, name + ".", x, True, False, training, bn_state, momentum, dp_masks.get(name), 1.0 - rates[bi])
        tap(name, x)
        bi += 1
    dense = batch_norm(sub, "norm.bn", x, training, bn_state, momentum)
    pooled = dense.mean(dim=(2, 3))
    return dense, pooled


# --------------------------------------------------------------------------------------
# Episode utilities and the MetaBaseline head
# --------------------------------------------------------------------------------------
def split_shot_query(data: Tensor, way: int, shot: int, query: int, ep_per_batch: int = 1):
    """utils/few_shot.py:4-10.  image ((e*way + c)*(shot+query) + j) is a shot if j < shot, else query
    number c*query + (j - shot) of episode e."""
    img = tuple(data.shape[1:])
    d = data.reshape(ep_per_batch, way, shot + query, *img)
    x_shot = d[:, :, :shot].contiguous()
    x_query = d[:, :, shot:].contiguous().reshape(ep_per_batch, way * query, *img)
    return x_shot, x_query


def make_nk_label(n: int, k: int, ep_per_batch: int = 1) -> Tensor:
    """utils/few_shot.py:13-16.  label = class id repeated k times, tiled per episode (int64)."""
    return torch.arange(n).repeat_interleave(k).repeat(ep_per_batch)


def l2_normalize(x: Tensor, eps: float = 1e-12) -> Tensor:
    """F.normalize(dim=-1): x / max(||x||_2, eps)."""
    return x / x.norm(dim=-1, keepdim=True).clamp_min(eps)


def compute_logits(feat: Tensor, proto: Tensor, metric: str = "dot", temp=1.0) -> Tensor:
    """utils/__init__.py:78-101 for 2-D and 3-D inputs, metrics 'dot' | 'cos' | 'sqr'."""
    assert feat.dim() == proto.dim()
    if metric == "cos":
        feat, proto = l2_normalize(feat), l2_normalize(proto)
        metric = "dot"
    if metric == "dot":
        logits = torch.matmul(feat, proto.transpose(-1, -2))
    elif metric == "sqr":
        logits = -((feat.unsqueeze(-2) - proto.unsqueeze(-3)) ** 2).sum(dim=-1)
    else:
        raise ValueError(metric)
    return logits * temp


def episode_head(feat_shot: Tensor, feat_query: Tensor, temp, method: str = "cos") -> Tensor:
    """meta_baseline.py:36-46.  feat_shot [E,way,shot,D], feat_query [E,Q,D] -> logits [E,Q,way]."""
    proto = feat_shot.mean(dim=-2)
    if method == "cos":
        return compute_logits(l2_normalize(feat_query), l2_normalize(proto), "dot", temp)
    return compute_logits(feat_query, proto, "sqr", temp)


def meta_baseline_forward(sd: SD, x_shot: Tensor, x_query: Tensor, method: str = "cos", training: bool = False,
                          bn_state=None, momentum=BN_MOMENTUM, drop_path_rate: float = 0.0, dp_masks=None) -> Tensor:
    """MetaBaseline.forward (meta_baseline.py:24-47): shots and queries go through the encoder as ONE
    batch (BN batch statistics span both in train mode), then the cosine head."""
    shot_shape, query_shape = x_shot.shape[:-3], x_query.shape[:-3]
    img = x_shot.shape[-3:]
    xs, xq = x_shot.reshape(-1, *img), x_query.reshape(-1, *img)
    _, feat = encoder_forward(sd, torch.cat([xs, xq], 0), "encoder.", training, bn_state, momentum,
                              drop_path_rate, dp_masks)
    fs, fq = feat[: xs.shape[0]], feat[xs.shape[0]:]
    return episode_head(fs.reshape(*shot_shape, -1), fq.reshape(*query_shape, -1), sd["temp"], method)


def cross_entropy(logits: Tensor, label: Tensor) -> Tensor:
    """F.cross_entropy mean reduction (test_few_shot.py:89, train_meta.py:169)."""
    lse = torch.logsumexp(logits, dim=-1)
    return (lse - logits.gather(1, label.view(-1, 1)).squeeze(1)).mean()


def compute_acc(logits: Tensor, label: Tensor) -> float:
    """utils/__init__.py:104-109."""
    return (logits.argmax(dim=1) == label).float().mean().item()


# --------------------------------------------------------------------------------------
# SUN local-supervision head
# --------------------------------------------------------------------------------------
def token_label_forward(sd: SD, x: Tensor, is_teacher: bool = False, training: bool = False, bn_state=None,
                        momentum=BN_MOMENTUM, drop_path_rate: float = 0.0, dp_masks=None):
    """TokenLabelOffline.forward (sun_meta_training/models/token_label.py:48-60).
    y_token is an NCHW *view* of NHWC memory ([B,n,5,5] with strides (25n, 1, 5n, n))."""
    dense, pooled = encoder_forward(sd, x, "encoder.", training, bn_state, momentum, drop_path_rate, dp_masks)
    name = "classifier" if is_teacher else "classifier_local"
    nhwc = dense.permute(0, 2, 3, 1)
    y_tok = torch.matmul(nhwc, sd[name + ".linear.weight"].t()) + sd[name + ".linear.bias"]
    y = torch.matmul(pooled, sd["classifier.linear.weight"].t()) + sd["classifier.linear.bias"]
    return y_tok.permute(0, 3, 1, 2), y, pooled


def generate_softlabel(logits: Tensor, smoothing: float = 0.1, k: int = 3, bp: int = 10) -> Tensor:
    """sun_meta_training/offline.py:57-76, bug-compatible.
    logits [B,n_cls,h,w] (teacher patch logits) -> soft labels [B*h*w, n_cls+1]:
      * foreground patches = the (h*w - bp) patches with the largest per-patch max logit;
      * a foreground row has `on` at its top-k class ids (un-normalised multi-hot), `off` elsewhere;
      * a background row has `on` at column **1**: the reference re-binds `c` to logits_max.size(1) == 1
        (offline.py:62) before `torch.full(..., c)` (offline.py:71), so the index is 1, not n_cls."""
    B, n_cls, h, w = logits.shape
    off = smoothing / n_cls
    on = 1.0 - smoothing + off
    flat = logits.permute(0, 2, 3, 1).reshape(B, h * w, n_cls)
    patch_max = flat.max(dim=-1).values                              # [B, hw]
    fg_idx = patch_max.topk(h * w - bp, dim=-1).indices
    fg = torch.zeros(B, h * w, dtype=torch.bool).scatter_(1, fg_idx, True).reshape(-1)
    rows = flat.reshape(-1, n_cls)
    top = rows.topk(k, dim=-1).indices
    out = torch.full((rows.shape[0], n_cls + 1), off, dtype=torch.float32)
    fg_rows = out.clone().scatter_(1, top, on)
    bg_rows = out.clone()
    bg_rows[:, 1] = on
    return torch.where(fg.view(-1, 1), fg_rows, bg_rows)


def soft_target_cross_entropy(x: Tensor, target: Tensor) -> Tensor:
    """SoftTargetCrossEntropy (offline.py:34-45): mean_rows(sum_c -t * log_softmax(x)); the target is
    tiled when x holds an integer multiple of its rows."""
    if x.shape[0] != target.shape[0]:
        target = target.repeat(x.shape[0] // target.shape[0], 1)
    return (-target * F.log_softmax(x, dim=-1)).sum(dim=-1).mean()


def sun_loss(logits_token: Tensor, logits: Tensor, label: Tensor, soft_label: Tensor) -> Tensor:
    """offline.py:270,296-300: CE(global) + 0.5 * token soft-CE (the 0.5 is hard-coded there)."""
    n = logits_token.shape[1]
    flat = logits_token.permute(0, 2, 3, 1).reshape(-1, n)
    return cross_entropy(logits, label) + 0.5 * soft_target_cross_entropy(flat, soft_label)


# --------------------------------------------------------------------------------------
# Fixtures: seeded synthetic episodes and BN calibration
# --------------------------------------------------------------------------------------
def make_episode_images(seed: int, way: int, per_class: int, structured: bool = True, noise: float = 0.5) -> Tensor:
    """[way*per_class, 3, 80, 80] fp32.  structured: x = p_c + noise*n (class prototypes p_c ~ N(0,1));
    otherwise iid N(0,1).  Image order = class-major, as CategoriesSampler batches are laid out."""
    g = torch.Generator().manual_seed(seed)
    if not structured:
        return torch.randn(way * per_class, 3, IMG, IMG, generator=g)
    protos = torch.randn(way, 1, 3, IMG, IMG, generator=g)
    n = torch.randn(way, per_class, 3, IMG, IMG, generator=g)
    return (protos + noise * n).reshape(way * per_class, 3, IMG, IMG)


def calibrate_bn(sd: SD, prefix: str = "encoder.", seed: int = 0, passes: int = 3) -> SD:
    """W1 fixture: running stats <- cumulative average (momentum=None) of `passes` train-mode forwards
    over 80 class-structured images (10 prototypes x 8)."""
    sd = dict(sd)
    for k in list(sd):
        if k.endswith("running_mean"):
            sd[k] = torch.zeros_like(sd[k])
        elif k.endswith("running_var"):
            sd[k] = torch.ones_like(sd[k])
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
    x = make_episode_images(seed, 10, 8)
    with torch.no_grad():
        for _ in range(passes):
            st = BNState()
            encoder_forward(sd, x, prefix, training=True, bn_state=st, momentum=None)
            for k, v in st.updates.items():
                sd[prefix + k] = v
    return sd


# --------------------------------------------------------------------------------------
# Input transform (test_phase/datasets/mini_imagenet.py:50-56): Resize((88,88)) -> CenterCrop(80) -> ToTensor -> Normalize
# --------------------------------------------------------------------------------------
def _pil_bilinear_coeffs(in_size: int, out_size: int):
    """PIL Resample.c precompute_coeffs (bilinear: support 1) + normalize_coeffs_8bpc (22 fractional bits)."""
    import numpy as np
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 1.0 * fscale
    ks = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([max(0.0, 1.0 - abs((x + xmin - center + 0.5) / fscale)) for x in range(xmax)], dtype=np.float64)
        w = w / w.sum()
        kk = np.where(w < 0, (-0.5 + w * (1 << 22)).astype(np.int64), (0.5 + w * (1 << 22)).astype(np.int64))
        ks.append((xmin, kk))
    return ks


def _pil_resample_axis(a, ks, axis):
    import numpy as np
    a = np.moveaxis(a, axis, 0).astype(np.int64)
    out = np.empty((len(ks),) + a.shape[1:], dtype=np.int64)
    for i, (xmin, kk) in enumerate(ks):
        out[i] = np.clip(((1 << 21) + np.tensordot(kk, a[xmin:xmin + len(kk)], axes=(0, 0))) >> 22, 0, 255)
    return np.moveaxis(out.astype(np.uint8), 0, axis)


def default_transform(img_u8, resize: int = 88, crop: int = 80, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)) -> Tensor:
    """One uint8 HWC image [S, S, 3] -> fp32 [3, crop, crop].  PIL resamples 8-bit images in fixed point, horizontally
    first, rounding / clipping to uint8 after each pass; CenterCrop offset = round((resize - crop) / 2); ToTensor = /255;
    Normalize = (x - mean) / std in fp32."""
    import numpy as np
    a = np.asarray(img_u8)
    ks = _pil_bilinear_coeffs(a.shape[0], resize)
    r = _pil_resample_axis(_pil_resample_axis(a, ks, 1), ks, 0)
    off = int(round((resize - crop) / 2.0))
    c = r[off:off + crop, off:off + crop]
    f = torch.from_numpy(np.ascontiguousarray(c)).permute(2, 0, 1).float().div(255)
    return (f - torch.tensor(mean).view(3, 1, 1)) / torch.tensor(std).view(3, 1, 1)


# --------------------------------------------------------------------------------------
# SUN-D head (meta_tuning_sun_d/Models/models/Network.py:48-81,109-175; emd_utils.py:65-76), evaluation / OpenCV solver
# --------------------------------------------------------------------------------------
def sund_weight_vector(A: Tensor, B: Tensor) -> Tensor:
    """Network.get_weight_vector: A [M,C,n,1], B [N,C,n,1] -> relu(<A node, mean-node of B>) + 1e-3, [M, N, n]."""
    Bm = B.mean(dim=(2, 3), keepdim=True)                               # adaptive_avg_pool2d(B, 1)
    comb = (A.unsqueeze(1) * Bm.unsqueeze(0)).sum(2)                    # [M, N, n, 1]
    return F.relu(comb.reshape(A.shape[0], B.shape[0], -1)) + 1e-3


def sund_similarity_map(proto: Tensor, query: Tensor) -> Tensor:
    """Network.get_similiarity_map (metric 'cosine') after normalize_feature ('center'): [Q, W, n_q, n_p]."""
    proto = proto - proto.mean(1).unsqueeze(1)
    query = query - query.mean(1).unsqueeze(1)
    p = proto.reshape(proto.shape[0], proto.shape[1], -1).permute(0, 2, 1)      # [W, n, C]
    q = query.reshape(query.shape[0], query.shape[1], -1).permute(0, 2, 1)      # [Q, n, C]
    return F.cosine_similarity(p[None, :, None, :, :], q[:, None, :, None, :], dim=-1)


def sund_emd_logits(proto: Tensor, query: Tensor, temperature: float = 12.5, return_flows: bool = False):
    """Network.emd_forward_1shot with solver 'opencv': cv2.EMD (OpenCV's transportation simplex, the third-party code the
    reference calls) on cost = 1 - similarity with node weights relu(w) + 1e-5 rescaled to sum n (emd_utils.py:65-76)."""
    import cv2
    w1 = sund_weight_vector(query, proto)                                # [Q, W, n]
    w2 = sund_weight_vector(proto, query)                                # [W, Q, n]
    sim = sund_similarity_map(proto, query)
    Q, W, n, _ = sim.shape
    logits = torch.zeros(Q, W)
    flows = torch.zeros(Q, W, n, n)
    for i in range(Q):
        for j in range(W):
            a = F.relu(w1[i, j]) + 1e-5
            b = F.relu(w2[j, i]) + 1e-5
            a = (a * (n / a.sum().item())).view(-1, 1).numpy()
            b = (b * (n / b.sum().item())).view(-1, 1).numpy()
            _, _, flow = cv2.EMD(a, b, cv2.DIST_USER, (1 - sim[i, j]).numpy())
            flows[i, j] = torch.from_numpy(flow)
            logits[i, j] = (sim[i, j] * flows[i, j]).sum() * (temperature / n)
    return (logits, flows) if return_flows else logits

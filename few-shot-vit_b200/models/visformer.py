"""Visformer encoder shim: the reference's module tree (identical state_dict names and shapes) whose
forward runs on the native sm_100a kernels through the C ABI (sunb_encoder_forward).

Reference: test_phase/models/visformer.py:291-462 (class), :482-487 (factory 'visformer_micro_80');
output variants: pooled only (test_phase :462), (dense, pooled) (sun_meta_training/models/visformer.py:464),
dense only (meta_tuning_sun_d/Models/models/visformer.py:461).

The sub-modules below are parameter containers (they keep `.train()/.eval()`, `freeze_bn`, optimizers and
`state_dict()/load_state_dict()` working); they never execute a torch forward.  Multi-GPU runs are one process per GPU
(`sunb200.dist`); wrapping the module in nn.DataParallel raises a clear error on the first replica forward.
"""
import math

import torch
import torch.nn as nn

from .models import register
from sunb200.engine import EncoderEngine
from sunb200.train import TrainEngine

IMG, STEM_CH, EMBED, DEPTH, HEADS, GROUPS = 80, 64, 256, (4, 2, 3), 6, 8


class _Holder(nn.Module):
    def forward(self, *a, **k):      # pragma: no cover - guard against accidental eager execution
        raise RuntimeError("sunb200 parameter container: the encoder forward runs in the native kernels")


def _conv(cin, cout, k, stride=1, padding=0, groups=1, bias=False):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=padding, groups=groups, bias=bias)


class LayerNorm(nn.Module):
    """Channel LayerNorm on NCHW tensors (reference wrapper visformer.py:109-115; unused by 'visformer_micro_80').
    Inference-only native kernel (sunb_layernorm_rows) over the NHWC view."""

    def __init__(self, dim):
        super().__init__()
        self.ln = nn.LayerNorm(dim)

    def forward(self, x):
        from sunb200 import native as N
        N.require_cuda(x)
        rows = x.permute(0, 2, 3, 1).contiguous().float()
        out = torch.empty_like(rows)
        M, Cc = rows.numel() // rows.shape[-1], rows.shape[-1]
        N.check(N.lib().sunb_layernorm_rows(rows.data_ptr(), self.ln.weight.detach().float().data_ptr(),
                                            self.ln.bias.detach().float().data_ptr(), out.data_ptr(), M, Cc,
                                            float(self.ln.eps), N.current_stream()), "sunb_layernorm_rows")
        return out.permute(0, 3, 1, 2)


class BatchNorm(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.bn = nn.BatchNorm2d(dim, eps=1e-5, momentum=0.1, track_running_stats=True)


class ConvBlock(_Holder):
    """Stem (visformer.py:202-239)."""

    def __init__(self, cin, hidden, planes):
        super().__init__()
        self.conv1 = _conv(cin, hidden, 3, stride=2, padding=1)
        self.bn1 = nn.BatchNorm2d(hidden)
        self.conv2 = _conv(hidden, planes, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = _conv(planes, planes, 3, padding=1)
        self.bn3 = nn.BatchNorm2d(planes)
        self.downsample = nn.Sequential(_conv(cin, planes, 3, stride=2, padding=1), nn.BatchNorm2d(planes))


class Mlp(_Holder):
    def __init__(self, dim, hidden, spatial_conv):
        super().__init__()
        if spatial_conv:
            hidden = dim * 2                      # visformer.py:136-141 with group >= 2
        self.conv1 = _conv(dim, hidden, 1)
        if spatial_conv:
            self.conv2 = _conv(hidden, hidden, 3, padding=1, groups=GROUPS)
        self.conv3 = _conv(hidden, dim, 1)


class Attention(_Holder):
    def __init__(self, dim, ratio=1.0):
        super().__init__()
        self.head_dim = round(dim // HEADS * ratio)
        self.qkv = _conv(dim, self.head_dim * HEADS * 3, 1)
        self.proj = _conv(self.head_dim * HEADS, dim, 1)


class Block(_Holder):
    def __init__(self, dim, attn, spatial_conv, drop_path):
        super().__init__()
        self.drop_prob = float(drop_path)
        if attn:
            self.norm1 = BatchNorm(dim)
            self.attn = Attention(dim)
        self.norm2 = BatchNorm(dim)
        self.mlp = Mlp(dim, dim * 4, spatial_conv)


class PatchEmbed(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.proj = nn.Conv2d(cin, cout, kernel_size=2, stride=2)
        self.norm = BatchNorm(cout)


class Visformer(nn.Module):
    """'visformer_micro_80' with output selection: output='pooled' | 'both' | 'dense'."""

    def __init__(self, drop_path_rate=0.0, output="pooled"):
        super().__init__()
        assert output in ("pooled", "both", "dense")
        self.output = output
        self.out_dim = EMBED * 2
        self.drop_path_rate = float(drop_path_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(DEPTH))]
        d1, d2, d3 = EMBED // 2, EMBED, EMBED * 2
        self.pos_embed1 = nn.Parameter(torch.zeros(1, d1, 20, 20))
        self.pos_embed2 = nn.Parameter(torch.zeros(1, d2, 10, 10))
        self.pos_embed3 = nn.Parameter(torch.zeros(1, d3, 5, 5))
        self.stem = ConvBlock(3, STEM_CH, d1)
        self.stage1 = nn.ModuleList([Block(d1, False, True, dpr[i]) for i in range(DEPTH[0])])
        self.patch_embed2 = PatchEmbed(d1, d2)
        self.stage2 = nn.ModuleList([Block(d2, True, False, dpr[DEPTH[0] + i]) for i in range(DEPTH[1])])
        self.patch_embed3 = PatchEmbed(d2, d3)
        self.stage3 = nn.ModuleList([Block(d3, True, False, dpr[DEPTH[0] + DEPTH[1] + i]) for i in range(DEPTH[2])])
        self.norm = BatchNorm(d3)
        self._reset_parameters()
        self._engine = EncoderEngine()
        self._train_engine = TrainEngine()

    def _reset_parameters(self):
        """Reference init distributions (visformer.py:398-422, conv_init=True)."""
        for p in (self.pos_embed1, self.pos_embed2, self.pos_embed3):
            nn.init.trunc_normal_(p, std=0.02, a=-2.0, b=2.0)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def _bn_in_eval(self):
        return all(not m.training for m in self.modules() if isinstance(m, nn.BatchNorm2d))

    def enable_data_parallel(self, group=None):
        """Multi-GPU meta-tuning (one process per GPU): all-reduce(mean) the encoder gradients inside the native backward
        pass, overlapped with the remaining backward kernels (replaces nn.DataParallel's reduce_add, train_meta.py:128-129).
        BatchNorm statistics stay per replica, as under DataParallel."""
        from sunb200.dist import GradComm
        self._train_engine.grad_comm = GradComm(group)
        return self

    def _drop_path_scales(self, batch, device, generator=None):
        """DropPath draws in the reference's forward order (visformer.py:89-97, 261-262): per block with rate > 0 one
        draw for the attention branch (stages 2/3) and one for the MLP; scale = floor(keep + U[0,1)) / keep."""
        rs = {}
        for stage, blocks in (("stage1", self.stage1), ("stage2", self.stage2), ("stage3", self.stage3)):
            for i, blk in enumerate(blocks):
                if blk.drop_prob <= 0.0:
                    continue
                keep = 1.0 - blk.drop_prob
                n = 1 if stage == "stage1" else 2
                rs[f"{stage}.{i}"] = [torch.floor(keep + torch.rand(batch, 1, 1, 1, device=device, generator=generator))
                                      .div_(keep).view(batch) for _ in range(n)]
        return rs

    # ------------------------------------------------------------------ packed-weight cache plumbing
    def _apply(self, fn, *args, **kwargs):
        # .cuda() / .to() / .float() replace the buffer tensors: drop the cached tensor list and the packed weights
        self._state_list = None
        self._engine.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._state_list = None
        self._engine.invalidate()
        return super().load_state_dict(*args, **kwargs)

    def invalidate_packed_weights(self):
        """Force a re-pack of the eval-mode weights on the next eval forward.  Needed only when parameters or BatchNorm
        statistics were changed by something torch's version counters do not see (CUDA-graph replays of a training step);
        the native train path and load_state_dict/.to() call it themselves."""
        self._engine.invalidate()

    def _state_tensors(self):
        """[(name, tensor)] of the 148 state entries; cached (tensor objects only change in _apply)."""
        if getattr(self, "_state_list", None) is None:
            self._state_list = list(self.state_dict(keep_vars=True).items())
        return self._state_list

    def _check_not_replica(self):
        if getattr(self, "_is_replica", False):
            raise RuntimeError(
                "sunb200: nn.DataParallel replicas are not supported (the native engines hold per-device packed weights and "
                "workspaces).  Run one process per GPU (torchrun) and shard episodes with sunb200.dist.shard_episodes; for "
                "meta-tuning call model.encoder.enable_data_parallel() (INTEGRATION.md section 4).")

    def features(self, x, taps=None, drop_path_scales=None):
        """dict(pooled [B,512] fp32, dense NHWC [B,5,5,512] fp32 | None, pooled_bf16, dense_bf16): both precisions of the
        final features as the last kernel emits them (the linear heads consume the bf16 copies directly)."""
        self._check_not_replica()
        bn_eval = self._bn_in_eval()
        want_dense = self.output != "pooled"
        if not self.training or (bn_eval and not torch.is_grad_enabled()):
            if not bn_eval:
                raise NotImplementedError("sunb200: eval-mode module with BatchNorm layers switched to train() is not supported")
            return self._engine.forward(self._state_tensors(), x, want_dense=want_dense, want_bf16=want_dense, taps=taps)
        # BatchNorm layers put in eval() by utils.freeze_bn use their running statistics inside the training step
        self._train_engine.frozen_bn = frozenset(n for n, m in self.named_modules()
                                                 if isinstance(m, nn.BatchNorm2d) and not m.training)
        rs = drop_path_scales if drop_path_scales is not None else self._drop_path_scales(x.shape[0], x.device)
        names = [n for n, _ in self.named_parameters()]
        params = [p for _, p in self.named_parameters()]
        pooled, dense, pooled16, dense16 = _EncoderTrainFn.apply(self, x, rs, names, *params)
        return {"pooled": pooled, "dense": dense, "pooled_bf16": pooled16, "dense_bf16": dense16}

    def forward(self, x, taps=None, drop_path_scales=None):
        out = self.features(x, taps=taps, drop_path_scales=drop_path_scales)
        pooled, dense = out["pooled"], out["dense"]
        if self.output == "pooled":
            return pooled
        dense = dense.permute(0, 3, 1, 2)              # NCHW view of NHWC memory (token_label.py:50 permutes it back)
        return (dense, pooled) if self.output == "both" else dense


class _EncoderTrainFn(torch.autograd.Function):
    """Train-mode encoder as one autograd node: forward and backward are the schedules in sunb200/train.py."""

    @staticmethod
    def forward(ctx, module, x, rs, names, *params):
        from sunb200 import native as N
        N.require_cuda(x)
        ctx.set_materialize_grads(False)
        P = {n: p.detach() for n, p in zip(names, params)}
        Bf = {n: b for n, b in module.named_buffers()}
        eng = module._train_engine
        # the native train path updates BatchNorm running statistics (and the caller's optimizer the weights) without
        # bumping torch's version counters in every case: the eval-mode packed weights are stale from here on
        module._engine.invalidate()
        pooled, dense, pooled16, dense16, c = eng.forward(P, Bf, x.detach().contiguous().float(), rs)
        ctx.eng, ctx.c, ctx.P, ctx.names = eng, c, P, names
        ctx.mark_non_differentiable(pooled16, dense16)
        return pooled, dense, pooled16, dense16

    @staticmethod
    def backward(ctx, dpooled, ddense, _dp16=None, _dd16=None):
        dp = dpooled.contiguous().float() if dpooled is not None else None
        dd = ddense.contiguous().float() if ddense is not None else None
        if dp is None and dd is None:
            return (None,) * (4 + len(ctx.names))
        G = ctx.eng.backward(ctx.P, ctx.c, dp, dd, comm=getattr(ctx.eng, "grad_comm", None))
        ctx.c = None
        return (None, None, None, None) + tuple(G[n] for n in ctx.names)


@register("visformer_micro_80")
def visformer_small_80(**kwargs):
    return Visformer(**kwargs)


@register("visformer")          # north_star spelling; SUN-D's --backbone value
def visformer(**kwargs):
    return Visformer(**kwargs)


@register("visformer_micro_80_dense")     # (dense, pooled) variant used by sun_meta_training
def visformer_small_80_dense(**kwargs):
    kwargs.setdefault("output", "both")
    return Visformer(**kwargs)

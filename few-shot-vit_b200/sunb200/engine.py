"""Host-side driver of the native encoder: packed-weight cache, workspace, forward call.

The reference mutates parameters behind the module's back (optimizer.step, load_state_dict, .cuda(), train/eval,
freeze_bn -- SURVEY.md 8b "packed-weight invalidation"), so the cache key is the tuple of every tensor's
(data_ptr, _version) plus the device; any change re-packs.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import native as N
from . import packing

TAP_SHAPES = {
    "stem": (20, 20, 128), "stage1.0": (20, 20, 128), "stage1.1": (20, 20, 128), "stage1.2": (20, 20, 128),
    "stage1.3": (20, 20, 128), "patch_embed2": (10, 10, 256), "stage2.0": (10, 10, 256), "stage2.1": (10, 10, 256),
    "patch_embed3": (5, 5, 512), "stage3.0": (5, 5, 512), "stage3.1": (5, 5, 512), "stage3.2": (5, 5, 512),
}


class EncoderEngine:
    """Eval-mode Visformer forward through libsunb200 (sunb_encoder_forward)."""

    def __init__(self):
        self._key = None
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._struct = None
        self._ws: Optional[torch.Tensor] = None
        self._ws_batch = 0

    @staticmethod
    def _state_key(tensors):
        return tuple((t.data_ptr(), t._version) for t in tensors)

    def invalidate(self) -> None:
        """Drop the packed weights: the next eval forward re-folds BatchNorm from the module's current state."""
        self._key = None

    def pack(self, state) -> None:
        """state: dict or [(name, tensor)] of the encoder state entries (reference names)."""
        items = list(state.items()) if isinstance(state, dict) else state
        key = self._state_key([t for _, t in items])
        if key == self._key:
            return
        with torch.no_grad():
            self._packed = packing.pack_encoder(dict(items))
        self._struct = packing.to_struct(self._packed)
        self._key = key

    def workspace(self, B: int, device) -> torch.Tensor:
        if self._ws is None or self._ws_batch < B or self._ws.device != device:
            nbytes = C.c_size_t(0)
            N.check(N.lib().sunb_encoder_workspace_bytes(B, C.byref(nbytes)), "sunb_encoder_workspace_bytes")
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
            self._ws_batch = B
        return self._ws

    def forward(self, state: Dict[str, torch.Tensor], x: torch.Tensor, want_dense: bool = False,
                want_bf16: bool = False, taps: Optional[Dict[str, torch.Tensor]] = None):
        """x fp32 NCHW [B,3,80,80] (CUDA).  Returns dict(pooled [B,512] fp32, dense NHWC [B,5,5,512] fp32 | None,
        dense_bf16, pooled_bf16).  `taps` (dict) is filled with bf16 NHWC layer boundaries when given."""
        N.require_cuda(x)
        if x.dtype != torch.float32 or x.dim() != 4 or tuple(x.shape[1:]) != (3, 80, 80):
            raise ValueError(f"encoder expects fp32 [B,3,80,80], got {x.dtype} {tuple(x.shape)}")
        x = x.contiguous()
        B = x.shape[0]
        self.pack(state)
        ws = self.workspace(B, x.device)
        pooled = torch.empty(B, 512, dtype=torch.float32, device=x.device)
        dense = torch.empty(B, 5, 5, 512, dtype=torch.float32, device=x.device) if want_dense else None
        dense16 = torch.empty(B, 5, 5, 512, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
        pooled16 = torch.empty(B, 512, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
        tstruct = None
        if taps is not None:
            tstruct = N.EncoderTaps()
            for name, shp in TAP_SHAPES.items():
                taps[name] = torch.empty(B, *shp, dtype=torch.bfloat16, device=x.device)
            tstruct.stem = taps["stem"].data_ptr()
            tstruct.patch_embed2 = taps["patch_embed2"].data_ptr()
            tstruct.patch_embed3 = taps["patch_embed3"].data_ptr()
            for i in range(4):
                tstruct.stage1[i] = taps[f"stage1.{i}"].data_ptr()
            for i in range(2):
                tstruct.stage2[i] = taps[f"stage2.{i}"].data_ptr()
            for i in range(3):
                tstruct.stage3[i] = taps[f"stage3.{i}"].data_ptr()
        st = N.lib().sunb_encoder_forward(C.byref(self._struct), x.data_ptr(), B, ws.data_ptr(), ws.numel(),
                                          pooled.data_ptr(), N.ptr(dense), N.ptr(dense16), N.ptr(pooled16),
                                          C.byref(tstruct) if tstruct is not None else None, N.current_stream())
        N.check(st, "sunb_encoder_forward")
        return {"pooled": pooled, "dense": dense, "dense_bf16": dense16, "pooled_bf16": pooled16}


def _temp_args(temp):
    if isinstance(temp, torch.Tensor) and temp.is_cuda:
        return temp.detach().float().reshape(1).contiguous(), 0.0
    return None, float(temp)


class _EpisodeLogitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fs, fq, temp, code):
        E, way, shot, D = fs.shape
        Q = fq.shape[1]
        fs, fq = fs.detach().contiguous().float(), fq.detach().contiguous().float()
        out = torch.empty(E, Q, way, dtype=torch.float32, device=fs.device)
        tdev, thost = _temp_args(temp)
        N.check(N.lib().sunb_episode_logits(fs.data_ptr(), fq.data_ptr(), out.data_ptr(), E, way, shot, Q, D, code,
                                            N.ptr(tdev), thost, N.current_stream()), "sunb_episode_logits")
        ctx.save_for_backward(fs, fq)
        ctx.temp, ctx.code = temp, code
        return out

    @staticmethod
    def backward(ctx, dlogits):
        fs, fq = ctx.saved_tensors
        E, way, shot, D = fs.shape
        Q = fq.shape[1]
        dl = dlogits.contiguous().float()
        dfs, dfq = torch.empty_like(fs), torch.empty_like(fq)
        temp = ctx.temp
        want_dt = isinstance(temp, torch.Tensor) and temp.requires_grad
        dt = torch.zeros((), dtype=torch.float32, device=fs.device) if want_dt else None
        tdev, thost = _temp_args(temp)
        N.check(N.lib().sunb_episode_logits_backward(fs.data_ptr(), fq.data_ptr(), dl.data_ptr(), dfs.data_ptr(),
                                                     dfq.data_ptr(), N.ptr(dt), E, way, shot, Q, D, ctx.code, N.ptr(tdev),
                                                     thost, N.current_stream()), "sunb_episode_logits_backward")
        return dfs, dfq, dt, None


def episode_logits(feat_shot: torch.Tensor, feat_query: torch.Tensor, temp, metric: str = "cos") -> torch.Tensor:
    """feat_shot [E,way,shot,D], feat_query [E,Q,D] (fp32 CUDA) -> logits [E,Q,way].
    Prototype mean + normalise + scaled dot product in one kernel (meta_baseline.py:36-46); differentiable
    (native backward) w.r.t. both feature tensors and a learnable `temp`."""
    N.require_cuda(feat_shot, feat_query)
    code = {"dot": 0, "cos": 1, "sqr": 2}[metric]
    return _EpisodeLogitsFn.apply(feat_shot, feat_query, temp, code)


def ce_and_acc(logits: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """[R,W] fp32 logits + int64 labels -> device tensor [mean CE, accuracy] without a host sync."""
    N.require_cuda(logits, label)
    l, y = logits.contiguous().float(), label.contiguous().long()
    out = torch.empty(2, dtype=torch.float32, device=l.device)
    N.check(N.lib().sunb_logits_ce_acc(l.data_ptr(), y.data_ptr(), l.shape[0], l.shape[1], out.data_ptr(),
                                       N.current_stream()), "sunb_logits_ce_acc")
    return out


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, label):
        l, y = logits.detach().contiguous().float(), label.contiguous().long()
        out = torch.empty(2, dtype=torch.float32, device=l.device)
        N.check(N.lib().sunb_logits_ce_acc(l.data_ptr(), y.data_ptr(), l.shape[0], l.shape[1], out.data_ptr(),
                                           N.current_stream()), "sunb_logits_ce_acc")
        ctx.save_for_backward(l, y)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        l, y = ctx.saved_tensors
        dl = torch.empty_like(l)
        gg = g.contiguous().float().reshape(1)
        N.check(N.lib().sunb_hard_ce_backward(l.data_ptr(), y.data_ptr(), l.shape[0], l.shape[1], gg.data_ptr(), 1.0,
                                              dl.data_ptr(), N.current_stream()), "sunb_hard_ce_backward")
        return dl, None


def cross_entropy(logits: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """F.cross_entropy (mean reduction) of [R,W] fp32 logits with a native forward and backward
    (reference: train_meta.py:169, offline.py:270)."""
    N.require_cuda(logits, label)
    return _CrossEntropy.apply(logits, label)


def generate_softlabel(logits: torch.Tensor, smoothing: float = 0.1, k: int = 3, bp: int = 10) -> torch.Tensor:
    """Teacher patch logits [B,n_cls,h,w] (any strides with a uniform pixel stride, e.g. the NHWC-backed view the
    token-label model returns) -> soft labels [B*h*w, n_cls+1] (sun_meta_training/offline.py:57-76)."""
    N.require_cuda(logits)
    logits = logits.float()
    B, n_cls, h, w = logits.shape
    sb, sc, sh, sw = logits.stride()
    if sh != sw * w:                       # no uniform pixel stride: make it NHWC-backed
        logits = logits.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        sb, sc, sh, sw = logits.stride()
    out = torch.empty(B * h * w, n_cls + 1, dtype=torch.float32, device=logits.device)
    N.check(N.lib().sunb_softlabel(logits.data_ptr(), sb, sc, sw, B, n_cls, h * w, k, bp, float(smoothing),
                                   out.data_ptr(), N.current_stream()), "sunb_softlabel")
    return out


class _SoftTargetCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target):
        N.require_cuda(x, target)
        x2, t2 = x.contiguous().float(), target.contiguous().float()
        R, Cc = x2.shape
        Rt = t2.shape[0]
        row = torch.empty(R, dtype=torch.float32, device=x2.device)
        loss = torch.empty((), dtype=torch.float32, device=x2.device)
        N.check(N.lib().sunb_soft_ce_forward(x2.data_ptr(), Cc, t2.data_ptr(), t2.shape[1], R, Rt, Cc, row.data_ptr(),
                                             loss.data_ptr(), N.current_stream()), "sunb_soft_ce_forward")
        ctx.save_for_backward(x2, t2)
        return loss

    @staticmethod
    def backward(ctx, g):
        x2, t2 = ctx.saved_tensors
        R, Cc = x2.shape
        dx = torch.empty_like(x2)
        gg = g.contiguous().float().reshape(1)
        N.check(N.lib().sunb_soft_ce_backward(x2.data_ptr(), Cc, t2.data_ptr(), t2.shape[1], R, t2.shape[0], Cc,
                                              gg.data_ptr(), 1.0, dx.data_ptr(), Cc, N.current_stream()),
                "sunb_soft_ce_backward")
        return dx, None


def soft_target_cross_entropy(x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """SoftTargetCrossEntropy (sun_meta_training/offline.py:34-45) with a native forward and backward."""
    return _SoftTargetCE.apply(x, target)

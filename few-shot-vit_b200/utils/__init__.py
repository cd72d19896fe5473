"""Drop-in `utils` package for the hot-path callers (reference: test_phase/utils/__init__.py).
compute_logits / compute_acc / linear run on the native kernels for CUDA tensors and raise on CPU tensors
(no CPU fallback).  The remaining helpers are host glue kept so the reference's scripts keep importing."""
import os
import shutil
import time

import torch
from torch.optim import Adam
from torch.optim.lr_scheduler import MultiStepLR

from . import few_shot  # noqa: F401
from sunb200 import engine as _engine
from sunb200 import native as _N
from sunb200.optim import FusedSGD, FusedAdamW, CosineLRScheduler, MultiStepLRScheduler  # noqa: F401

_log_path = None


def set_log_path(path):
    global _log_path
    _log_path = path


def log(obj, filename="log.txt"):
    print(obj)
    if _log_path is not None:
        with open(os.path.join(_log_path, filename), "a") as f:
            print(obj, file=f)


class Averager:
    """Running mean."""

    def __init__(self):
        self.n, self.v = 0.0, 0.0

    def add(self, v, n=1.0):
        self.v = (self.v * self.n + v * n) / (self.n + n)
        self.n += n

    def item(self):
        return self.v


class Timer:
    def __init__(self):
        self.v = time.time()

    def s(self):
        self.v = time.time()

    def t(self):
        return time.time() - self.v


def set_gpu(gpu):
    print("set gpu:", gpu)
    os.environ["CUDA_VISIBLE_DEVICES"] = gpu


def ensure_path(path, remove=True):
    base = os.path.basename(path.rstrip("/"))
    if os.path.exists(path):
        if remove and (base.startswith("_") or input(f"{path} exists, remove? ([y]/n): ") != "n"):
            shutil.rmtree(path)
            os.makedirs(path)
    else:
        os.makedirs(path)


def time_str(t):
    if t >= 3600:
        return f"{t / 3600:.1f}h"
    if t >= 60:
        return f"{t / 60:.1f}m"
    return f"{t:.1f}s"


def compute_logits(feat, proto, metric="dot", temp=1.0):
    """[E,Q,D] x [E,W,D] (or 2-D [Q,D] x [W,D]) -> logits * temp; metric 'dot' | 'cos' | 'sqr'
    (reference: utils/__init__.py:78-101)."""
    assert feat.dim() == proto.dim()
    if feat.dim() == 2:
        return _engine.episode_logits(proto.unsqueeze(0).unsqueeze(2), feat.unsqueeze(0), temp, metric)[0]
    return _engine.episode_logits(proto.unsqueeze(2), feat, temp, metric)


def compute_acc(logits, label, reduction="mean"):
    """argmax(dim=1) == label (reference: utils/__init__.py:104-109)."""
    if reduction == "none":
        return (torch.argmax(logits, dim=1) == label).float().detach()
    return _engine.ce_and_acc(logits.detach(), label)[1].item()


def _pad8(n):
    return (n + 7) // 8 * 8


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM with a native backward: dX = dY W (sunb_gemm), dW = dY^T X (sunb_wgrad, split-K on
    the tensor cores), db = column sums of dY (sunb_colstats).  bf16 operands, fp32 accumulation and outputs."""

    @staticmethod
    def forward(ctx, x, weight, bias, x_bf16):
        import ctypes as C
        K = x.shape[-1]
        n = weight.shape[0]
        a = x_bf16 if x_bf16 is not None else x.detach().reshape(-1, K).to(torch.bfloat16)
        a = a.reshape(-1, K).contiguous()
        w = weight.detach().to(torch.bfloat16).contiguous()
        out = torch.empty(a.shape[0], n, dtype=torch.float32, device=x.device)
        d = _N.GemmDesc()
        d.M, d.N, d.K, d.taps, d.groups = a.shape[0], n, K, 1, 1
        d.A, d.lda, d.Wt, d.ldw = a.data_ptr(), K, w.data_ptr(), K
        b32 = bias.detach().float().contiguous() if bias is not None else None
        d.bias, d.bias_mod, d.bias_ld = _N.ptr(b32), 1, 0
        d.rows_per_img = 1
        d.out_f32, d.ldc_f32 = out.data_ptr(), n
        _N.check(_N.lib().sunb_gemm(C.byref(d), 0, _N.current_stream()), "sunb_gemm(linear)")
        ctx.save_for_backward(a, w)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return out.reshape(*x.shape[:-1], n)

    @staticmethod
    def backward(ctx, dy):
        import ctypes as C
        a, w = ctx.saved_tensors
        M, K = a.shape
        n = w.shape[0]
        ldn = _pad8(n)                                   # TMA rows are multiples of 16 bytes: pad N = 65 -> 72 with zeros
        st = _N.current_stream()
        dyb = torch.zeros(M, ldn, dtype=torch.bfloat16, device=a.device)
        dyb[:, :n] = dy.reshape(M, n)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wt = torch.zeros(K, ldn, dtype=torch.bfloat16, device=a.device)      # dgrad operand: [K rows][n], reduction over n
            wt[:, :n] = w.t()
            dx = torch.empty(M, K, dtype=torch.float32, device=a.device)
            d = _N.GemmDesc()
            d.M, d.N, d.K, d.taps, d.groups = M, K, n, 1, 1
            d.A, d.lda, d.Wt, d.ldw = dyb.data_ptr(), ldn, wt.data_ptr(), ldn
            d.bias_mod, d.rows_per_img = 1, 1
            d.out_f32, d.ldc_f32 = dx.data_ptr(), K
            _N.check(_N.lib().sunb_gemm(C.byref(d), 0, st), "sunb_gemm(linear dgrad)")
            dx = dx.reshape(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(n, K, dtype=torch.float32, device=a.device)
            g = _N.WgradDesc()
            g.P, g.Ma, g.Nb, g.Ca, g.Cb = M, n, K, n, K
            g.groups, g.taps = 1, 1
            g.dY, g.ldy, g.X, g.ldx = dyb.data_ptr(), ldn, a.data_ptr(), K
            g.out, g.ldo, g.ksplit = dw.data_ptr(), K, 0
            _N.check(_N.lib().sunb_wgrad(C.byref(g), st), "sunb_wgrad(linear)")
        if ctx.has_bias and ctx.needs_input_grad[2]:
            sums = torch.zeros(2, ldn, dtype=torch.float32, device=a.device)
            _N.check(_N.lib().sunb_colstats(dyb.data_ptr(), ldn, None, 0, M, ldn, sums[0].data_ptr(), sums[1].data_ptr(), st),
                     "sunb_colstats(linear bias)")
            db = sums[0, :n].clone()
        return dx, dw, db, None


def linear(x, weight, bias=None, x_bf16=None):
    """nn.Linear on the tcgen05 GEMM: x [..., K] fp32 -> [..., N] fp32 (bf16 operands, fp32 accumulate), differentiable
    w.r.t. x, weight and bias through native kernels.  `x_bf16` (optional) is an existing bf16 copy of x with the same shape
    (the encoder's final kernel emits one), which saves the cast."""
    _N.require_cuda(x, weight)
    return _LinearFn.apply(x, weight, bias, x_bf16)


def compute_n_params(model, return_str=True):
    tot = sum(p.numel() for p in model.parameters())
    if not return_str:
        return tot
    return f"{tot / 1e6:.1f}M" if tot >= 1e6 else f"{tot / 1e3:.1f}K"


def make_optimizer(params, name, lr, weight_decay=None, milestones=None, gamma=0.1):
    """SGD(momentum 0.9) / Adam + optional MultiStepLR (reference: utils/__init__.py:128-139).  'adamw' (used by
    sun_meta_training/offline.py:229 directly) is accepted too."""
    wd = 0.0 if weight_decay is None else weight_decay
    if name == "sgd":
        opt = FusedSGD(params, lr, momentum=0.9, weight_decay=wd)      # one native launch over all 86 tensors
    elif name == "adam":
        opt = Adam(params, lr, weight_decay=wd)
    elif name == "adamw":
        opt = FusedAdamW(params, lr, weight_decay=wd)
    else:
        raise ValueError(name)
    sched = MultiStepLR(opt, milestones, gamma=gamma) if milestones else None
    return opt, sched


def visualize_dataset(dataset, name, writer, n_samples=16):
    """Image grid of random samples for tensorboard (reference: utils/__init__.py:142-147); host glue."""
    import numpy as np
    demo = [dataset.convert_raw(dataset[i][0]) for i in np.random.choice(len(dataset), n_samples)]
    writer.add_images("visualize_" + name, torch.stack(demo))
    writer.flush()


def freeze_bn(model):
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()

"""Drop-in `utils` package for the hot-path callers (reference: test_phase/utils/__init__.py).
compute_logits / compute_acc / linear run on the native kernels for CUDA tensors and raise on CPU tensors
(no CPU fallback).  The remaining helpers are host glue kept so the reference's scripts keep importing."""
import os
import shutil
import time

import torch
from torch.optim import SGD, Adam
from torch.optim.lr_scheduler import MultiStepLR

from . import few_shot  # noqa: F401
from sunb200 import engine as _engine
from sunb200 import native as _N

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


def linear(x, weight, bias=None):
    """nn.Linear forward on the tcgen05 GEMM: x [..., K] fp32 -> [..., N] fp32 (bf16 operands, fp32 accumulate)."""
    import ctypes as C
    _N.require_cuda(x, weight)
    lead, K = x.shape[:-1], x.shape[-1]
    a = x.reshape(-1, K).to(torch.bfloat16).contiguous()
    w = weight.detach().to(torch.bfloat16).contiguous()
    n = w.shape[0]
    out = torch.empty(a.shape[0], n, dtype=torch.float32, device=x.device)
    d = _N.GemmDesc()
    d.M, d.N, d.K, d.taps, d.groups = a.shape[0], n, K, 1, 1
    d.A, d.lda, d.Wt, d.ldw = a.data_ptr(), K, w.data_ptr(), K
    b32 = bias.detach().float().contiguous() if bias is not None else None
    d.bias, d.bias_mod, d.bias_ld = _N.ptr(b32), 1, 0
    d.rows_per_img = 1
    d.out_f32, d.ldc_f32 = out.data_ptr(), n
    _N.check(_N.lib().sunb_gemm(C.byref(d), 0, _N.current_stream()), "sunb_gemm(linear)")
    return out.reshape(*lead, n)


def compute_n_params(model, return_str=True):
    tot = sum(p.numel() for p in model.parameters())
    if not return_str:
        return tot
    return f"{tot / 1e6:.1f}M" if tot >= 1e6 else f"{tot / 1e3:.1f}K"


def make_optimizer(params, name, lr, weight_decay=None, milestones=None, gamma=0.1):
    """SGD(momentum 0.9) / Adam + optional MultiStepLR (reference: utils/__init__.py:128-139)."""
    wd = 0.0 if weight_decay is None else weight_decay
    if name == "sgd":
        opt = SGD(params, lr, momentum=0.9, weight_decay=wd)
    elif name == "adam":
        opt = Adam(params, lr, weight_decay=wd)
    else:
        raise ValueError(name)
    sched = MultiStepLR(opt, milestones, gamma=gamma) if milestones else None
    return opt, sched


def freeze_bn(model):
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()

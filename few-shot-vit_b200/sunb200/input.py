"""On-device input path: the evaluation / meta-tuning transform of the reference's datasets and the episode gather.

Reference: test_phase/datasets/mini_imagenet.py:50-56 (`default_transform` = Resize((88, 88)) -> CenterCrop(80) -> ToTensor ->
Normalize) applied per image on DataLoader workers, and datasets/samplers.py:21-35 (CategoriesSampler) for the episode
indices.  Here the uint8 image store (84 x 84 x 3 per image, as the miniImageNet pickles hold it) lives in HBM, an episode
batch is a gather index, and one kernel (csrc/input.cu) does gather + PIL-exact resize + crop + normalise -> fp32 NCHW.
"""
from __future__ import annotations

import functools
import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import native as N

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 22          # PIL Resample.c: 32 - 8 - 2
KMAX = 3


@functools.lru_cache(maxsize=8)
def resize_tables(in_size: int = 84, resize: int = 88, crop: int = 80) -> Tuple[np.ndarray, np.ndarray]:
    """(first source index int32 [crop], fixed-point weights int32 [crop, 3]) of every output position kept by the centre
    crop: PIL's precompute_coeffs (bilinear, support 1 * max(scale, 1)) followed by normalize_coeffs_8bpc."""
    scale = in_size / resize
    fscale = max(scale, 1.0)
    support = 1.0 * fscale
    off = int(round((resize - crop) / 2.0))            # torchvision CenterCrop
    tmin = np.zeros(crop, dtype=np.int32)
    tk = np.zeros((crop, KMAX), dtype=np.int32)
    for o in range(crop):
        xx = o + off
        center = (xx + 0.5) * scale
        ss = 1.0 / fscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.array([max(0.0, 1.0 - abs((x + xmin - center + 0.5) * ss)) for x in range(xmax)], dtype=np.float64)
        w = w / w.sum()
        k = [int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in w]
        while k and k[-1] == 0:
            k.pop()
        lead = 0
        while lead < len(k) and k[lead] == 0:
            lead += 1
        k = k[lead:]
        if len(k) > KMAX:
            raise ValueError("resize_tables: more than 3 taps per output (down-sampling is not supported by the kernel)")
        tmin[o] = xmin + lead
        tk[o, :len(k)] = k
    return tmin, tk


class _Tables:
    def __init__(self, device, in_size, resize, crop, mean, std):
        tmin, tk = resize_tables(in_size, resize, crop)
        self.tmin = torch.from_numpy(tmin).to(device)
        self.tk = torch.from_numpy(tk).to(device).contiguous()
        self.mean_std = torch.tensor(list(mean) + list(std), dtype=torch.float32, device=device)
        self.in_size, self.crop = in_size, crop


_cache = {}


def preprocess_u8(data_u8: torch.Tensor, index: Optional[torch.Tensor] = None, resize: int = 88, crop: int = 80,
                  mean: Sequence[float] = MEAN, std: Sequence[float] = STD) -> torch.Tensor:
    """data_u8: CUDA uint8 [N, S, S, 3] (HWC).  index: optional int64 [n] gather index (host or device; a sampler batch is a
    host tensor: 8 bytes of H2D per image).  Returns fp32 [n, 3, crop, crop]."""
    N.require_cuda(data_u8)
    if data_u8.dtype != torch.uint8 or data_u8.dim() != 4 or data_u8.shape[-1] != 3 or data_u8.shape[1] != data_u8.shape[2]:
        raise ValueError(f"preprocess_u8 expects uint8 [N, S, S, 3], got {data_u8.dtype} {tuple(data_u8.shape)}")
    data_u8 = data_u8.contiguous()
    key = (str(data_u8.device), data_u8.shape[1], resize, crop, tuple(mean), tuple(std))
    if key not in _cache:
        _cache[key] = _Tables(data_u8.device, data_u8.shape[1], resize, crop, mean, std)
    t = _cache[key]
    idx = None
    if index is not None:
        if index.numel() and (int(index.min()) < 0 or int(index.max()) >= data_u8.shape[0]):     # host tensors: no device sync
            raise IndexError("preprocess_u8: gather index out of range")
        idx = index.to(device=data_u8.device, dtype=torch.int64, non_blocking=True).contiguous()
    n = data_u8.shape[0] if idx is None else idx.numel()
    out = torch.empty(n, 3, crop, crop, dtype=torch.float32, device=data_u8.device)
    if n:
        N.check(N.lib().sunb_preprocess_u8(data_u8.data_ptr(), N.ptr(idx), n, t.in_size, t.crop, t.tmin.data_ptr(), t.tk.data_ptr(),
                                           t.mean_std.data_ptr(), out.data_ptr(), N.current_stream()), "sunb_preprocess_u8")
    return out


class DeviceImageStore:
    """uint8 image store resident in HBM (e.g. the 12 000 test images of miniImageNet = 254 MB) + labels on the host.
    `batch(indices)` = what a DataLoader over the reference dataset yields for a CategoriesSampler batch: (fp32 images
    [n, 3, 80, 80] on the device, int64 labels)."""

    def __init__(self, data_u8, labels, device="cuda"):
        data = torch.as_tensor(np.asarray(data_u8)) if not isinstance(data_u8, torch.Tensor) else data_u8
        self.data = data.to(device=device, dtype=torch.uint8).contiguous()
        labels = np.asarray(labels)
        self.label = list((labels - labels.min()).tolist())       # mini_imagenet.py:40-41
        self.n_classes = max(self.label) + 1
        self._label_t = torch.as_tensor(self.label, dtype=torch.int64)

    def __len__(self):
        return self.data.shape[0]

    def batch(self, indices, unchecked: bool = False):
        idx = torch.as_tensor(indices, dtype=torch.int64)
        if unchecked:                      # indices produced by our own sampler: skip the range check (it syncs)
            dev_idx = idx.to(self.data.device, non_blocking=True)
            t = _cache.get((str(self.data.device), self.data.shape[1], 88, 80, MEAN, STD))
            if t is None:
                return preprocess_u8(self.data, dev_idx), self._label_t[idx]
            out = torch.empty(idx.numel(), 3, t.crop, t.crop, dtype=torch.float32, device=self.data.device)
            N.check(N.lib().sunb_preprocess_u8(self.data.data_ptr(), dev_idx.data_ptr(), idx.numel(), t.in_size, t.crop,
                                               t.tmin.data_ptr(), t.tk.data_ptr(), t.mean_std.data_ptr(), out.data_ptr(),
                                               N.current_stream()), "sunb_preprocess_u8")
            return out, self._label_t[idx]
        return preprocess_u8(self.data, idx), self._label_t[idx]

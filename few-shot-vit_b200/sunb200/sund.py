"""SUN-D: DeepEMD-style patch-to-patch head on the native Visformer encoder, evaluation path (1-shot).

Reference: meta_tuning_sun_d/Models/models/Network.py (`DeepEMD`): `encode` (:177-187, grid mode: 13 crops per image, each
encoded and average-pooled to one 512-d node), `emd_forward_1shot` (:67-81), `get_emd_distance` with the OpenCV solver
(:109-124, one cv2.EMD call and one device->host copy per (query, class) pair), eval loop eval.py:83-98.
Here the node features come from the native encoder (pooled output per crop) and the whole head -- node weights, centring,
cosine similarity map and the transportation problem -- is one kernel launch per episode (csrc/emd.cu).
The k > 1 shot path of the reference fine-tunes prototypes through the qpth QP solver (`get_sfc`); qpth is not part of
this image and that path is out of scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import native as N


def emd_logits(proto: torch.Tensor, query: torch.Tensor, temperature: float = 12.5, return_flows: bool = False):
    """proto [W, C, n(,1)], query [Q, C, n(,1)] fp32 CUDA node features (Network.py layout) -> logits [Q, W]
    (= Network.emd_forward_1shot in eval mode / solver 'opencv', metric 'cosine', norm 'center')."""
    N.require_cuda(proto, query)
    p = proto.reshape(proto.shape[0], proto.shape[1], -1).permute(0, 2, 1).contiguous().float()     # [W, n, C]
    q = query.reshape(query.shape[0], query.shape[1], -1).permute(0, 2, 1).contiguous().float()     # [Q, n, C]
    W, n, D = p.shape
    Q = q.shape[0]
    logits = torch.empty(Q, W, dtype=torch.float32, device=p.device)
    flows = torch.empty(Q, W, n, n, dtype=torch.float32, device=p.device) if return_flows else None
    N.check(N.lib().sunb_emd_head(p.data_ptr(), q.data_ptr(), logits.data_ptr(), N.ptr(flows), W, Q, n, D, float(temperature),
                                  N.current_stream()), "sunb_emd_head")
    return (logits, flows) if return_flows else logits


class DeepEMD(nn.Module):
    """Evaluation-mode drop-in for the reference's DeepEMD(args) with backbone 'visformer' and the grid-crop loader:
    mode 'encoder': forward(x [N, n_crops, 3, 80, 80]) -> node features [N, 512, n_crops, 1];
    mode 'meta'   : forward((proto [1, W, 512, n, 1], query [Q, 512, n, 1])) -> logits [Q, W]."""

    def __init__(self, temperature: float = 12.5, mode: str = "meta"):
        super().__init__()
        import models
        self.encoder = models.make("visformer_micro_80")          # pooled output == avg-pool of the dense map (Network.py:182-183)
        self.dim = 512
        self.temperature = float(temperature)
        self.mode = mode

    def encode(self, x):
        if x.dim() == 5:                                           # batch of image crops (Network.py:179-187)
            num_data, num_patch = x.shape[:2]
            feat = self.encoder(x.reshape(-1, *x.shape[2:]))       # [N * n, 512]
            return feat.reshape(num_data, num_patch, self.dim).permute(0, 2, 1).unsqueeze(-1)
        return self.encoder(x).unsqueeze(-1).unsqueeze(-1)

    def emd_forward_1shot(self, proto, query):
        return emd_logits(proto.squeeze(0), query, self.temperature)

    def forward(self, inp):
        if self.mode == "encoder":
            return self.encode(inp)
        if self.mode == "meta":
            support, query = inp
            return self.emd_forward_1shot(support, query)
        raise ValueError("Unknown mode")

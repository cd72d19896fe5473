"""MetaBaseline episode model on the native kernels (reference: test_phase/models/meta_baseline.py:10-47)."""
import torch
import torch.nn as nn

import models
from .models import register
from sunb200 import engine


@register("meta-baseline")
class MetaBaseline(nn.Module):
    """forward(x_shot[E,way,shot,3,80,80], x_query[E,Q,3,80,80]) -> logits [E,Q,way] (fp32)."""

    def __init__(self, encoder, encoder_args={}, method="cos", temp=10., temp_learnable=True):
        super().__init__()
        self.encoder = models.make(encoder, **encoder_args)
        if method not in ("cos", "sqr"):
            raise ValueError(f"unknown method '{method}'")
        self.method = method
        self.temp = nn.Parameter(torch.tensor(float(temp))) if temp_learnable else temp

    def forward(self, x_shot, x_query):
        lead_s, lead_q = x_shot.shape[:-3], x_query.shape[:-3]
        img = x_shot.shape[-3:]
        n_s = x_shot.numel() // (img[0] * img[1] * img[2])
        # one encoder batch [shots ; queries]: in train mode the BN statistics span both (meta_baseline.py:31)
        batch = torch.cat([x_shot.reshape(-1, *img), x_query.reshape(-1, *img)], dim=0)
        feat = self.encoder(batch)
        if isinstance(feat, tuple):
            feat = feat[-1]
        f_shot = feat[:n_s].reshape(*lead_s, -1)
        f_query = feat[n_s:].reshape(*lead_q, -1)
        if f_shot.dim() == 3:                      # single episode without the leading E dimension
            f_shot, f_query = f_shot.unsqueeze(0), f_query.unsqueeze(0)
            return engine.episode_logits(f_shot, f_query, self.temp, self.method)[0]
        return engine.episode_logits(f_shot, f_query, self.temp, self.method)

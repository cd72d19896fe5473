"""SUN meta-training step and validation on the native kernels (reference: sun_meta_training/offline.py).

`sun_meta_training_step` restates the body of the reference's hot loop (offline.py:263-303):
    logits_token, logits, token = student(strong)              # train mode, DropPath, BN batch statistics  (:269)
    cls_loss = CE(logits, label)                               # (:270)
    no_grad: teacher(weak, True) -> patch logits -> generate_softlabel(k = tl_soft_k, bp = bg_token_num)  (:288-290)
    token_loss = SoftTargetCrossEntropy(student patch logits, soft labels)                                (:296-297)
    loss = cls_loss + 0.5 * token_loss                         # the 0.5 is hard-coded there (:300)
    zero_grad / backward / optimizer.step                      # AdamW(lr 5e-4 * bs/512, wd 0.05) (:228-229, :301-303)
The student runs on the train engine (sunb200/train.py), the frozen teacher on the eval engine (BatchNorm folded), the
linear heads on the tcgen05 GEMM with native dgrad / wgrad (utils.linear), CE / soft-label / soft-CE on their kernels, and
the update in one fused AdamW launch (sunb200/optim.py).  `build_optimizer` gives the reference's optimizer + schedule.

`sun_meta_validate` restates the validation body (offline.py:321-336): queries and shots go through the model
separately, prototypes = mean over shots, cosine logits with temperature 10.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import engine
from .optim import CosineLRScheduler, FusedAdamW


def build_optimizer(model: torch.nn.Module, batch_size: int, lr: float = 5e-4, weight_decay: float = 0.05,
                    max_epoch: int = 800, warmup: int = 5, warmup_lr: float = 1e-6):
    """AdamW + cosine schedule of offline.py:228-231 (configs/offline_tl_visformer_k5_800epoch.yaml:21-27)."""
    opt = FusedAdamW(model.parameters(), betas=(0.9, 0.999), eps=1e-8, lr=float(lr) * (batch_size / 512),
                     weight_decay=float(weight_decay))
    sched = CosineLRScheduler(opt, t_initial=max_epoch, cycle_decay=0.1, warmup_t=int(warmup), warmup_lr_init=float(warmup_lr))
    return opt, sched


def sun_losses(student, teacher, strong, weak, label, tl_soft_k: int = 5, bg_token_num: int = 10,
               soft_label: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Forward part of the step.  `soft_label` overrides the teacher's pseudo labels (parity tests inject the
    reference's)."""
    logits_token, logits, _ = student(strong)
    cls_loss = engine.cross_entropy(logits, label)
    if soft_label is None:
        with torch.no_grad():
            logits_token_t, _, _ = teacher(weak, True)
            soft_label = engine.generate_softlabel(logits_token_t, k=tl_soft_k, bp=bg_token_num)
    c1 = logits_token.shape[1]
    flat = logits_token.permute(0, 2, 3, 1).reshape(-1, c1)          # a view: y_token is NHWC-backed (token_label.py:56)
    token_loss = engine.soft_target_cross_entropy(flat, soft_label)
    return {"loss": cls_loss + 0.5 * token_loss, "cls_loss": cls_loss, "token_loss": token_loss, "logits": logits,
            "logits_token": logits_token, "soft_label": soft_label}


def sun_meta_training_step(student, teacher, strong, weak, label, optimizer, tl_soft_k: int = 5, bg_token_num: int = 10,
                           soft_label: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    out = sun_losses(student, teacher, strong, weak, label, tl_soft_k, bg_token_num, soft_label)
    optimizer.zero_grad(set_to_none=True)
    out["loss"].backward()
    optimizer.step()
    return out


@torch.no_grad()
def sun_meta_validate(model, x_shot, x_query, temp: float = 10.0):
    """x_shot [E,way,shot,3,80,80], x_query [E,Q,3,80,80] -> (cosine logits [E*Q, way], device tensor [mean CE, acc])
    against the labels the caller supplies through `utils.few_shot.make_nk_label` (offline.py:321-336)."""
    E, way, shot = x_shot.shape[:3]
    Q = x_query.shape[1]
    img = x_shot.shape[-3:]
    _, _, q_tok = model(x_query.reshape(-1, *img))
    _, _, s_tok = model(x_shot.reshape(-1, *img))
    q_tok = q_tok.reshape(E, Q, -1)
    s_tok = s_tok.reshape(E, way, shot, -1)
    return engine.episode_logits(s_tok, q_tok, temp, "cos").reshape(-1, way)

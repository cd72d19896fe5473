"""SUN local-supervision model 'token-label' (reference: sun_meta_training/models/token_label.py:36-60)."""
import torch.nn as nn

import models
import utils
from .models import register


@register("token-label")
class TokenLabelOffline(nn.Module):
    """forward(x, is_teacher=False) -> (y_token [B,n,5,5] NCHW view of NHWC memory, y [B,n_cls], pooled [B,512]).
    Student: classifier_local (n_cls+1 outputs incl. the background column); teacher: the global classifier per patch."""

    def __init__(self, encoder, encoder_args, classifier, classifier_args):
        super().__init__()
        encoder_args = dict(encoder_args, output="both")
        self.encoder = models.make(encoder, **encoder_args)
        base = dict(classifier_args, in_dim=self.encoder.out_dim)
        local = dict(base, n_classes=int(classifier_args["n_classes"]) + 1)
        self.classifier = models.make(classifier, **base)
        self.classifier_local = models.make(classifier, **local)

    def forward(self, x, is_teacher=False):
        head = self.classifier if is_teacher else self.classifier_local
        if hasattr(self.encoder, "features"):                   # native encoder: fp32 + bf16 copies of the final features
            f = self.encoder.features(x)
            tokens, pooled = f["dense"], f["pooled"]            # dense is NHWC-contiguous [B,5,5,512]
            y_token = _apply_head(head, tokens, f["dense_bf16"]).permute(0, 3, 1, 2)
            return y_token, _apply_head(self.classifier, pooled, f["pooled_bf16"]), pooled
        dense, pooled = self.encoder(x)
        tokens = dense.permute(0, 2, 3, 1)                      # NHWC, contiguous
        y_token = head(tokens).permute(0, 3, 1, 2)
        return y_token, self.classifier(pooled), pooled


def _apply_head(head, x, x_bf16):
    """LinearClassifier heads take the bf16 copy the encoder's last kernel already wrote (no re-cast of the activations)."""
    lin = getattr(head, "linear", None)
    if isinstance(lin, nn.Linear):
        return utils.linear(x, lin.weight, lin.bias, x_bf16=x_bf16)
    return head(x)

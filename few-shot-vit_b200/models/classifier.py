"""Classifier heads of the SUN pre-/meta-training stage (reference: sun_meta_training/models/classifier.py:11-54)."""
import math

import torch
import torch.nn as nn

import models
import utils
from .models import register


@register("linear-classifier")
class LinearClassifier(nn.Module):
    def __init__(self, in_dim, n_classes):
        super().__init__()
        self.linear = nn.Linear(in_dim, n_classes)

    def forward(self, x):
        return utils.linear(x, self.linear.weight, self.linear.bias)


@register("nn-classifier")
class NNClassifier(nn.Module):
    def __init__(self, in_dim, n_classes, metric="cos", temp=None):
        super().__init__()
        self.proto = nn.Parameter(torch.empty(n_classes, in_dim))
        nn.init.kaiming_uniform_(self.proto, a=math.sqrt(5))
        if temp is None:
            temp = nn.Parameter(torch.tensor(10.)) if metric == "cos" else 1.0
        self.metric, self.temp = metric, temp

    def forward(self, x):
        return utils.compute_logits(x, self.proto, self.metric, self.temp)


@register("classifier")
class Classifier(nn.Module):
    """encoder -> pooled feature -> classifier."""

    def __init__(self, encoder, encoder_args, classifier, classifier_args):
        super().__init__()
        self.encoder = models.make(encoder, **encoder_args)
        classifier_args = dict(classifier_args, in_dim=self.encoder.out_dim)
        self.classifier = models.make(classifier, **classifier_args)

    def forward(self, x):
        feat = self.encoder(x)
        if isinstance(feat, tuple):
            feat = feat[-1]
        return self.classifier(feat)

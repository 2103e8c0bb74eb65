"""FPN top-down merge (reference backbones/fpn_top_down.py:6-30): 1x1 lateral convs (no bias) on the
coarsest-first feature list, running sum with bilinear up-sampling, one 3x3 merge conv at the end."""
import torch.nn as nn
import torch.nn.functional as F


class FPNTopDown(nn.Module):
    def __init__(self, pyramid_channels, feature_channel):
        super().__init__()
        self.reduction_layers = nn.ModuleList(
            nn.Conv2d(c, feature_channel, 1, bias=False) for c in pyramid_channels)
        self.merge_layer = nn.Conv2d(feature_channel, feature_channel, 3, padding=1, bias=False)

    def upsample_add(self, x, y):
        return F.interpolate(x, size=tuple(y.shape[2:]), mode='bilinear') + y

    def forward(self, pyramid_features):
        merged = None
        for level, lateral in zip(pyramid_features, self.reduction_layers):
            level = lateral(level)
            merged = level if merged is None else self.upsample_add(merged, level)
        return self.merge_layer(merged)

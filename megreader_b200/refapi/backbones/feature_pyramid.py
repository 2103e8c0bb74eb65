"""bottom_up trunk -> top_down merge on the reversed feature tuple (reference backbones/feature_pyramid.py:4-14)."""
import torch.nn as nn


class FeaturePyramid(nn.Module):
    def __init__(self, bottom_up, top_down):
        super().__init__()
        self.bottom_up = bottom_up
        self.top_down = top_down

    def forward(self, feature):
        return self.top_down(self.bottom_up(feature)[::-1])

"""Pyramid pooling head (reference backbones/ppm.py:6-44).

Input: the list of trunk features; only the last (N, fc_dim, H, W) is used.  Four adaptive-average-pooled
copies (1,2,3,6 bins) go through 1x1 conv(->512)+BN+ReLU, are bilinearly resized back (align_corners=False)
and concatenated with the input; `conv_last` = 3x3 conv(->512)+BN+ReLU+Dropout2d(0.1)+1x1 conv(->inner).
`cbr_deepsup` is constructed (state-dict parity) but unused in forward, as in the reference."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .base import conv3x3_bn_relu


class PPMDeepsup(nn.Module):
    def __init__(self, inner_channels=256, fc_dim=2048, pool_scales=(1, 2, 3, 6)):
        super().__init__()
        self.ppm = nn.ModuleList(
            nn.Sequential(nn.AdaptiveAvgPool2d(bins), nn.Conv2d(fc_dim, 512, 1, bias=False),
                          nn.BatchNorm2d(512), nn.ReLU(inplace=True))
            for bins in pool_scales)
        self.cbr_deepsup = conv3x3_bn_relu(fc_dim // 2, fc_dim // 4, 1)
        self.conv_last = nn.Sequential(
            nn.Conv2d(fc_dim + 512 * len(pool_scales), 512, 3, padding=1, bias=False),
            nn.BatchNorm2d(512), nn.ReLU(inplace=True), nn.Dropout2d(0.1),
            nn.Conv2d(512, inner_channels, 1))

    def forward(self, conv_out, segSize=None):
        top = conv_out[-1]
        size = top.shape[2:]
        pooled = [F.interpolate(branch(top), tuple(size), mode='bilinear', align_corners=False)
                  for branch in self.ppm]
        return self.conv_last(torch.cat([top] + pooled, 1))

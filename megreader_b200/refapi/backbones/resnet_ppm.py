"""`resnet50dilated_ppm` (reference backbones/resnet_ppm.py:9-13): Sequential(ResnetDilated(resnet50, 8), PPMDeepsup)."""
import torch.nn as nn

from .resnet import resnet50
from .resnet_dilated import ResnetDilated
from .ppm import PPMDeepsup


def resnet50dilated_ppm(resnet_pretrained=False, **kwargs):
    trunk = ResnetDilated(resnet50(pretrained=resnet_pretrained), dilate_scale=8)
    return nn.Sequential(trunk, PPMDeepsup(**kwargs))

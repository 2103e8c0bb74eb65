"""ResNet + FPN factories (reference backbones/resnet_fpn.py:6-38).  `Resnet34FPN` builds a resnet50 trunk in
the reference (resnet_fpn.py:14) — kept, since checkpoints trained there depend on it."""
from . import resnet as _resnet
from .fpn_top_down import FPNTopDown
from .feature_pyramid import FeaturePyramid

_TRUNK = {'Resnet18FPN': ('resnet18', [512, 256, 128, 64]), 'Resnet34FPN': ('resnet50', [2048, 1024, 512, 256]),
          'Resnet50FPN': ('resnet50', [2048, 1024, 512, 256]), 'Resnet101FPN': ('resnet101', [2048, 1024, 512, 256]),
          'Resnet152FPN': ('resnet152', [2048, 1024, 512, 256])}


def _factory(name):
    trunk, channels = _TRUNK[name]

    def make(resnet_pretrained=True):
        return FeaturePyramid(getattr(_resnet, trunk)(pretrained=resnet_pretrained), FPNTopDown(channels, 256))
    make.__name__ = make.__qualname__ = name
    return make


Resnet18FPN, Resnet34FPN, Resnet50FPN, Resnet101FPN, Resnet152FPN = (_factory(n) for n in _TRUNK)

"""Dilated view of a ResNet trunk (reference backbones/resnet_dilated.py:5-69).

`dilate_scale=8` removes the stride of layer3/layer4 and dilates their 3x3 convs by 2/4 (the conv that lost
its stride gets half of that); `dilate_scale=16` does the same for layer4 only with 2.  The wrapped
trunk's own modules are re-registered under the same names, so state-dict keys are those of ResNet minus
avgpool/fc/smooth.  `forward` returns the list [x2, x3, x4, x5] (or x5 alone)."""
import torch.nn as nn

_PARTS = ('conv1', 'bn1', 'relu1', 'conv2', 'bn2', 'relu2', 'conv3', 'bn3', 'relu3', 'maxpool',
          'layer1', 'layer2', 'layer3', 'layer4')


def _pair(v):
    return (v, v)


class ResnetDilated(nn.Module):
    def __init__(self, orig_resnet, dilate_scale=8):
        super().__init__()
        plan = {8: (('layer3', 2), ('layer4', 4)), 16: (('layer4', 2),)}.get(dilate_scale, ())
        for stage, rate in plan:
            for m in getattr(orig_resnet, stage).modules():
                self._nostride_dilate(m, rate)
        for name in _PARTS:
            setattr(self, name, getattr(orig_resnet, name))

    def _nostride_dilate(self, m, dilate):
        if 'Conv' not in type(m).__name__:             # name test as in the reference: DCN modules match too
            return
        three = tuple(m.kernel_size) == (3, 3)
        if tuple(m.stride) == (2, 2):
            m.stride = (1, 1)
            if three:
                m.dilation = m.padding = _pair(dilate // 2)
        elif three:
            m.dilation = m.padding = _pair(dilate)

    def forward(self, x, return_feature_maps=True):
        for i in (1, 2, 3):
            x = getattr(self, 'relu%d' % i)(getattr(self, 'bn%d' % i)(getattr(self, 'conv%d' % i)(x)))
        x = self.maxpool(x)
        conv_out = []
        for stage in (self.layer1, self.layer2, self.layer3, self.layer4):
            x = stage(x)
            conv_out.append(x)
        return conv_out if return_feature_maps else x

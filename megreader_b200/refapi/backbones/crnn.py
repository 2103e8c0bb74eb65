"""CRNN backbone with the reference's surface (backbones/crnn.py:4-63): `crnn_backbone(imgH=32, nc=3, nclass=37,
nh=256)` -> module whose `.cnn` is a 7-block Sequential with the same state-dict keys:
cnn.{0,1,3,5}.0.0 = conv+ReLU (+pool), cnn.{2,4,6}.{0,1} = conv + BatchNorm (NO ReLU, App. B3.1)."""
import torch.nn as nn

from megreader_b200 import crnn_engine


class CRNN(nn.Module):
    KERNELS = [3, 3, 3, 3, 3, 3, 2]
    PADDINGS = [1, 1, 1, 1, 1, 1, 0]
    CHANNELS = [64, 128, 256, 256, 512, 512, 512]
    # pooling after block i: (kernel, stride, padding) — backbones/crnn.py:18-35
    POOLS = {0: ((2, 2), (2, 2), (0, 0)), 1: ((2, 2), (2, 2), (0, 0)),
             3: ((2, 2), (2, 1), (0, 1)), 5: ((2, 2), (2, 1), (0, 1))}
    BN_BLOCKS = (2, 4, 6)

    def __init__(self, imgH, nc, nclass, nh):
        super().__init__()
        assert imgH % 16 == 0, 'imgH has to be a multiple of 16'
        blocks = []
        cin = nc
        for i, cout in enumerate(self.CHANNELS):
            conv = nn.Conv2d(cin, cout, self.KERNELS[i], 1, self.PADDINGS[i])
            if i in self.BN_BLOCKS:
                block = nn.Sequential(conv, nn.BatchNorm2d(cout))
            else:
                block = nn.Sequential(conv, nn.ReLU())
                if i in self.POOLS:
                    k, s, p = self.POOLS[i]
                    block = nn.Sequential(block, nn.MaxPool2d(k, s, p))
                else:                      # unreachable with the reference's layout; kept for clarity
                    block = nn.Sequential(block)
            blocks.append(block)
            cin = cout
        self.cnn = nn.Sequential(*blocks)

    def forward(self, input):
        return crnn_engine.backbone_forward(self, input)


def crnn_backbone(imgH=32, nc=3, nclass=37, nh=256):
    return CRNN(imgH, nc, nclass, nh)

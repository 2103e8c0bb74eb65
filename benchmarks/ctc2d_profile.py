"""Launch each 2D-CTC kernel a few times at a saturating batch (for ncu).  python benchmarks/ctc2d_profile.py [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.ctc2d_micro import make  # noqa: E402
from megreader_b200 import ctc2d  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda:0")
lp, tg, il, tl = make(N, dev)
go = 1.0 / tl.float()
for _ in range(2):
    nll, la = ctc2d.ctc2d_forward(lp, tg, il, tl, 0, 0.0)
    g = ctc2d.ctc2d_backward(go, lp, tg, il, tl, nll, la, 0)
    nll2, gfac = ctc2d.ctc2d_forward_train(lp, tg, il, tl, 0)
    g2 = ctc2d.ctc2d_backward_apply(go, lp, gfac)
torch.cuda.synchronize()
print("ok", float(nll.sum()), float(g.abs().sum()), float(g2.abs().sum()))

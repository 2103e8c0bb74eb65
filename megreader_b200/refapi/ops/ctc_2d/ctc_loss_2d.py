# mirrors ops/ctc_2d/ctc_loss_2d.py:1-37
from megreader_b200.ctc2d import CTCLoss2DFunction, ctc_loss_2d  # noqa: F401
from . import ctc_2d_csrc  # noqa: F401

# mirrors ops/__init__.py:1
from .ctc_2d.ctc_loss_2d import CTCLoss2DFunction, ctc_loss_2d  # noqa: F401

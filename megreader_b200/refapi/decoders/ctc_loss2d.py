# mirrors decoders/ctc_loss2d.py (class CTCLoss2D; CTC2DLoss is the name BASELINE.json uses — SURVEY.md D1)
from megreader_b200.ctc2d import CTC2DLoss, CTCLoss2D  # noqa: F401

# mirrors decoders/__init__.py for the hot-path heads
from .crnn import CRNNDecoder  # noqa: F401
from .ctc_decoder2d import CTCDecoder2D  # noqa: F401
from .ctc_loss2d import CTCLoss2D, CTC2DLoss  # noqa: F401

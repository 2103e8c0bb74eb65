# mirrors decoders/__init__.py for the recognition heads (detection heads are out of scope)
from .attention_decoder import AttentionDecoder  # noqa: F401
from .ctc_decoder import CTCDecoder  # noqa: F401
from .crnn import CRNNDecoder  # noqa: F401
from .ctc_decoder2d import CTCDecoder2D  # noqa: F401
from .ctc_loss2d import CTCLoss2D, CTC2DLoss  # noqa: F401
from .east import EASTDecoder  # noqa: F401

"""Execution engine behind the CRNN surfaces (refapi/backbones/crnn.py, refapi/decoders/crnn.py).

The nn.Modules only own parameters (reference names / shapes); the arithmetic is dispatched
here.  Stage status (DESIGN.md §kernels keeps this table current):
    conv stack / BN / pools   : library (ATen -> cuDNN) in this revision
    BiLSTM + Linear           : library (ATen -> cuDNN / cuBLAS) in this revision
    log_softmax + 1D CTC loss : megreader_b200 CUDA (csrc/ctc2d.cu, H = 1 path)
CUDA only: there is no CPU execution path.
"""
import torch
import torch.nn.functional as F

from . import ctc1d


def _require_cuda(t, what):
    if not t.is_cuda:
        raise NotImplementedError("megreader_b200.%s: CUDA tensors only (no CPU fallback)" % what)


def backbone_forward(module, x):
    """backbones/crnn.py:57-59."""
    _require_cuda(x, "crnn_backbone")
    return module.cnn(x)


def bilstm_forward(module, x):
    """decoders/crnn.py:16-24: (T, N, nIn) -> LSTM -> Linear -> (T, N, nOut)."""
    _require_cuda(x, "BidirectionalLSTM")
    recurrent, _ = module.rnn(x)
    T, b, h = recurrent.size()
    return module.embedding(recurrent.view(T * b, h)).view(T, b, -1)


def decoder_forward(module, feature, targets=None, lengths=None, train=False):
    """decoders/crnn.py:80-104."""
    _require_cuda(feature, "CRNNDecoder")
    b, c, h, w = feature.size()
    if h > 1:
        feature = module.fpn2rnn(feature)
        b, c, h, w = feature.size()
    assert h == 1, "the height of conv must be 1"
    seq = feature.squeeze(2).permute(2, 0, 1)          # (W, N, C)
    for r in module.rnn:
        r.rnn.flatten_parameters()
    pred = module.rnn(seq)                             # (T, N, classes)
    if train:
        T = pred.size(0)
        pred_size = torch.full((b,), T, dtype=torch.int64, device=pred.device)
        if module.loss_func == 'pytorch':
            loss, lp = ctc1d.ctc_loss_from_logits(pred.float(), targets, pred_size, lengths, blank=0,
                                                  zero_infinity=True, reduction="mean")
        else:   # decoders/ctc_loss.py:118-121: per-sample nll / target_length, no zero_infinity
            nll, lp = ctc1d.ctc_loss_from_logits(pred.float(), targets, pred_size, lengths, blank=0,
                                                 zero_infinity=False, reduction="none")
            loss = nll / lengths.to(nll.device).to(nll.dtype)
        return loss, lp.to(torch.float64)              # decoders/crnn.py:96 hands float64 log-probs back
    pred = pred.permute(1, 2, 0).unsqueeze(2)
    return F.softmax(pred, dim=1)

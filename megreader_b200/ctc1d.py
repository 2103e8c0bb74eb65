"""1D CTC head: the `log_softmax(dim=2) -> nn.CTCLoss(zero_infinity=True)` pair of the reference's CRNNDecoder
(decoders/crnn.py:47-48,95-99), fused on the device: megreader_b200/csrc/ctc2d.cu (H = 1 DP) through the C-ABI.
The python restatement in the reference (decoders/ctc_loss.py:65-122) defines the same arithmetic."""
import torch
from torch.autograd import Function

from . import _lib
from . import ctc2d as _ctc2d


def _stream():
    return torch.cuda.current_stream().cuda_stream


class LogSoftmaxCTCFunction(Function):
    """(logits[T,N,C] fp32, targets[N,S], input_lengths[N], target_lengths[N]) -> (nll[N], log_probs[T,N,C]).
    nll follows torch's 'none' reduction with zero_infinity applied; log_probs is returned non-differentiable
    (the reference only hands it to visualisers, structure/model.py:178-180)."""

    @staticmethod
    def forward(ctx, logits, targets, input_lengths, target_lengths, blank, zero_infinity):
        if not logits.is_cuda:
            raise NotImplementedError("megreader_b200: the CTC head runs on CUDA only (no CPU fallback)")
        logits = logits.contiguous().float()
        T, N, C = logits.shape
        dev = logits.device
        tg = targets.to(device=dev, dtype=torch.int64)
        if tg.dim() != 2:
            raise RuntimeError("targets must be [N, S] (the concatenated form is not implemented, decoders/ctc_loss.py:71)")
        il = input_lengths.to(device=dev, dtype=torch.int64).contiguous()
        tl = target_lengths.to(device=dev, dtype=torch.int64).contiguous()
        lp = torch.empty_like(logits)
        nll = torch.empty((N,), dtype=torch.float32, device=dev)
        gfac = torch.empty_like(logits)
        L = _lib.lib()
        with torch.cuda.device(dev):
            _lib.check(L.mr_log_softmax_rows_f32(logits.data_ptr(), T * N, C, lp.data_ptr(), _stream()), "log_softmax")
            _lib.check(L.mr_ctc1d_forward_train_f32(
                lp.data_ptr(), tg.data_ptr(), il.data_ptr(), tl.data_ptr(), T, N, C, tg.size(1), tg.stride(0),
                tg.stride(1), blank, int(zero_infinity), int(_ctc2d.FAST_MATH), nll.data_ptr(), gfac.data_ptr(),
                _stream()), "ctc1d_forward")
        if zero_infinity:
            nll = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll)
        ctx.save_for_backward(lp, gfac)
        ctx.mark_non_differentiable(lp)
        return nll, lp

    @staticmethod
    def backward(ctx, grad_nll, _grad_lp):
        lp, gfac = ctx.saved_tensors
        T, N, C = lp.shape
        scale = grad_nll.contiguous().float()
        grad = torch.empty_like(lp)
        with torch.cuda.device(lp.device):
            _lib.check(_lib.lib().mr_ctc1d_backward_logits_f32(lp.data_ptr(), gfac.data_ptr(), scale.data_ptr(), T, N,
                                                               C, grad.data_ptr(), _stream()), "ctc1d_backward")
        return grad, None, None, None, None, None


def ctc_loss_from_logits(logits, targets, input_lengths, target_lengths, blank=0, zero_infinity=True,
                         reduction="mean"):
    """== nn.CTCLoss(blank, reduction, zero_infinity)(log_softmax(logits, 2), ...) ; also returns log_probs."""
    nll, lp = LogSoftmaxCTCFunction.apply(logits, targets, input_lengths, target_lengths, blank, zero_infinity)
    if reduction == "mean":   # aten: mean over the batch of nll / clamp(target_length, 1)
        tl = target_lengths.to(nll.device).clamp(min=1).to(nll.dtype)
        return (nll / tl).mean(), lp
    if reduction == "sum":
        return nll.sum(), lp
    return nll, lp

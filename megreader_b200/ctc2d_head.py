"""Fused epilogue of the 2D-CTC head: conv-branch logits -> log_probs (T,H,N,C) in one pass, and the whole
`CTCDecoder2D` training tail (decoders/ctc_decoder2d.py:37-49) as one autograd node whose backward never
materialises d(log_probs).  CUDA only (csrc/ctc2d_head.cu through the C-ABI); no CPU fallback."""
import torch
from torch.autograd import Function

from . import _lib
from . import ctc2d as _ctc2d

TINY = float(torch.finfo(torch.float32).tiny)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_inputs(mask_logits, cls_logits):
    if not (mask_logits.is_cuda and cls_logits.is_cuda):
        raise NotImplementedError("megreader_b200: the 2D-CTC head epilogue runs on CUDA only (no CPU fallback)")
    if cls_logits.dim() != 4 or mask_logits.dim() != 4 or mask_logits.size(1) != 1 or \
            mask_logits.shape[0] != cls_logits.shape[0] or mask_logits.shape[2:] != cls_logits.shape[2:]:
        raise RuntimeError("expected mask_logits [N,1,H,W] and classify_logits [N,C,H,W]")


def head_forward(mask_logits, cls_logits, tiny=TINY):
    N, C, H, W = cls_logits.shape
    lp = torch.empty((W, H, N, C), dtype=torch.float32, device=cls_logits.device)
    with torch.cuda.device(cls_logits.device):
        _lib.check(_lib.lib().mr_ctc2d_head_fwd_f32(mask_logits.data_ptr(), cls_logits.data_ptr(), N, C, H, W, tiny,
                                                    lp.data_ptr(), _stream()), "ctc2d_head_fwd")
    return lp


def head_backward(mask_logits, cls_logits, grad_lp=None, gfac=None, grad_out=None, tiny=TINY):
    N, C, H, W = cls_logits.shape
    dcls = torch.empty_like(cls_logits)
    dmask = torch.empty_like(mask_logits)
    with torch.cuda.device(cls_logits.device):
        _lib.check(_lib.lib().mr_ctc2d_head_bwd_f32(
            mask_logits.data_ptr(), cls_logits.data_ptr(), grad_lp.data_ptr() if grad_lp is not None else None,
            gfac.data_ptr() if gfac is not None else None, grad_out.data_ptr() if grad_out is not None else None,
            (grad_out.stride(0) if grad_out.dim() else 0) if grad_out is not None else 0, N, C, H, W, tiny,
            dcls.data_ptr(), dmask.data_ptr(), _stream()), "ctc2d_head_bwd")
    return dmask, dcls


class HeadLogProbsFunction(Function):
    """(mask_logits [N,1,H,W], classify_logits [N,C,H,W]) -> log_probs [W,H,N,C]; differentiable in both inputs."""

    @staticmethod
    def forward(ctx, mask_logits, cls_logits, tiny):
        _check_inputs(mask_logits, cls_logits)
        m, z = mask_logits.contiguous().float(), cls_logits.contiguous().float()
        ctx.save_for_backward(m, z)
        ctx.tiny = tiny
        return head_forward(m, z, tiny)

    @staticmethod
    def backward(ctx, grad_lp):
        m, z = ctx.saved_tensors
        dmask, dcls = head_backward(m, z, grad_lp=grad_lp.contiguous().float(), tiny=ctx.tiny)
        return dmask, dcls, None


def head_log_probs(mask_logits, cls_logits, tiny=TINY):
    return HeadLogProbsFunction.apply(mask_logits, cls_logits, tiny)


class HeadLossFunction(Function):
    """logits -> (nll [N], log_probs [W,H,N,C]): head epilogue + 2D-CTC training forward; backward goes from grad_nll
    straight to the logits' gradients through the per-(t,class) factor (include/megreader_b200.h).  log_probs is a
    non-differentiable by-product (the reference hands it to the representer, structure/model.py:178-180)."""

    @staticmethod
    def forward(ctx, mask_logits, cls_logits, targets, input_lengths, target_lengths, blank, tiny):
        _check_inputs(mask_logits, cls_logits)
        m, z = mask_logits.contiguous().float(), cls_logits.contiguous().float()
        lp = head_forward(m, z, tiny)
        nll, gfac = _ctc2d.ctc2d_forward_train(lp, targets, input_lengths, target_lengths, blank)
        ctx.save_for_backward(m, z, gfac)
        ctx.tiny = tiny
        ctx.mark_non_differentiable(lp)
        return nll, lp

    @staticmethod
    def backward(ctx, grad_nll, _grad_lp):
        m, z, gfac = ctx.saved_tensors
        dmask, dcls = head_backward(m, z, gfac=gfac, grad_out=grad_nll.contiguous().float(), tiny=ctx.tiny)
        return dmask, dcls, None, None, None, None, None


def head_loss(mask_logits, cls_logits, targets, input_lengths, target_lengths, blank=0, tiny=TINY):
    return HeadLossFunction.apply(mask_logits, cls_logits, targets, input_lengths, target_lengths, blank, tiny)

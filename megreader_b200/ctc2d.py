"""Host side of the 2D-CTC op: same names, arguments and error behaviour as the reference's
pybind module `ops.ctc_2d.ctc_2d_csrc` (ops/ctc_2d/csrc/ctc2d.cpp:3-6, ctc2d.h:7-43) and its
autograd wrapper `CTCLoss2DFunction` / `ctc_loss_2d` (ops/ctc_2d/ctc_loss_2d.py:7-37), plus the
(mask, classify) module `CTCLoss2D` (decoders/ctc_loss2d.py:8-154; alias CTC2DLoss).

All arithmetic happens in megreader_b200/csrc/ctc2d.cu through the C-ABI.
"""
import os

import torch
from torch.autograd import Function

from . import _lib

# ex2/lg2.approx inside the kernels (f32).  Parity tests cover both settings.
FAST_MATH = os.environ.get("MEGREADER_B200_FAST_MATH", "1") != "0"
# True: CTCLoss2DFunction keeps the reference's exact data flow (forward materialises log_alpha,
# backward = ctc2d_backward).  False (default): the training pair (no log_alpha in HBM).
CONTRACT_PATH = os.environ.get("MEGREADER_B200_CTC2D_CONTRACT", "0") == "1"


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_inputs(log_probs, targets, input_lengths, target_lengths, blank):
    if not log_probs.is_cuda:
        raise RuntimeError("Not implemented on the CPU")                      # ctc2d.h:20
    if log_probs.dim() != 4:
        raise RuntimeError("log_probs must be [T, H, N, C]")
    if not log_probs.is_contiguous():
        raise RuntimeError("log_probs tensor has to be contiguous")           # ctc2d_cuda.cu:35
    T, H, N, C = log_probs.shape
    if not (0 <= blank < C):
        raise RuntimeError("blank must be in label range")                    # ctc2d_cuda.cu:40
    if input_lengths.size(0) != N:
        raise RuntimeError("input_lengths must be of size batch_size")        # ctc2d_cuda.cu:41
    if target_lengths.size(0) != N:
        raise RuntimeError("target_lengths must be of size batch_size")       # ctc2d_cuda.cu:42
    for name, t in (("targets", targets), ("input_lengths", input_lengths), ("target_lengths", target_lengths)):
        # the reference reads these through data<int64_t>() on the device (ctc2d_cuda_kernel.cu:240-242)
        if t.dtype != torch.int64:
            raise RuntimeError("expected scalar type Long for %s but got %s" % (name, t.dtype))
        if not t.is_cuda:
            raise RuntimeError("%s must be a CUDA tensor" % name)
    if targets.dim() != 2 or targets.size(0) != N:
        raise RuntimeError("targets must be [N, S]")
    S = targets.size(1)
    if 2 * S + 1 > 1024:
        raise RuntimeError("max target length out of range, got %d, must less than 1024" % S)  # kernel.cu:220
    if log_probs.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("megreader_b200 ctc2d: dtype %s not supported (float32/float64 only)" % log_probs.dtype)
    return T, H, N, C, S


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def ctc2d_forward(log_probs, targets, input_lengths, target_lengths, BLANK, TINY=0.0):
    """-> (neg_log_likelihood[N], log_alpha[N,T,H,2S+1]).  ctc2d.h:7-21.  TINY is unused (as in the reference)."""
    T, H, N, C, S = _check_inputs(log_probs, targets, input_lengths, target_lengths, BLANK)
    il, tl = _c(input_lengths), _c(target_lengths)
    nll = torch.empty((N,), dtype=log_probs.dtype, device=log_probs.device)
    log_alpha = torch.empty((N, T, H, 2 * S + 1), dtype=log_probs.dtype, device=log_probs.device)
    L = _lib.lib()
    fn = L.mr_ctc2d_forward_f32 if log_probs.dtype == torch.float32 else L.mr_ctc2d_forward_f64
    with torch.cuda.device(log_probs.device):
        _lib.check(fn(log_probs.data_ptr(), targets.data_ptr(), il.data_ptr(), tl.data_ptr(), T, H, N, C, S,
                      targets.stride(0), targets.stride(1), BLANK, int(FAST_MATH),
                      nll.data_ptr(), log_alpha.data_ptr(), _stream()), "ctc2d_forward")
    return nll, log_alpha


def ctc2d_backward(grad_out, log_probs, targets, input_lengths, target_lengths, neg_log_likelihood, log_alpha, BLANK):
    """-> grad[T,H,N,C].  ctc2d.h:24-43.  log_alpha / nll are accepted and not read (see ctc2d.cu header)."""
    T, H, N, C, S = _check_inputs(log_probs, targets, input_lengths, target_lengths, BLANK)
    il, tl = _c(input_lengths), _c(target_lengths)
    if grad_out.dtype != log_probs.dtype:
        grad_out = grad_out.to(log_probs.dtype)
    grad = torch.empty_like(log_probs)
    L = _lib.lib()
    fn = L.mr_ctc2d_backward_f32 if log_probs.dtype == torch.float32 else L.mr_ctc2d_backward_f64
    with torch.cuda.device(log_probs.device):
        _lib.check(fn(grad_out.data_ptr(), grad_out.stride(0) if grad_out.dim() else 0, log_probs.data_ptr(),
                      targets.data_ptr(), il.data_ptr(), tl.data_ptr(),
                      neg_log_likelihood.data_ptr() if neg_log_likelihood is not None else None,
                      log_alpha.data_ptr() if log_alpha is not None else None,
                      T, H, N, C, S, targets.stride(0), targets.stride(1), BLANK, int(FAST_MATH),
                      grad.data_ptr(), _stream()), "ctc2d_backward")
    return grad


def ctc2d_forward_train(log_probs, targets, input_lengths, target_lengths, BLANK):
    """-> (nll[N], gfac[T,N,C]) — training forward without log_alpha (float32)."""
    T, H, N, C, S = _check_inputs(log_probs, targets, input_lengths, target_lengths, BLANK)
    il, tl = _c(input_lengths), _c(target_lengths)
    nll = torch.empty((N,), dtype=log_probs.dtype, device=log_probs.device)
    gfac = torch.empty((T, N, C), dtype=log_probs.dtype, device=log_probs.device)
    with torch.cuda.device(log_probs.device):
        _lib.check(_lib.lib().mr_ctc2d_forward_train_f32(
            log_probs.data_ptr(), targets.data_ptr(), il.data_ptr(), tl.data_ptr(), T, H, N, C, S,
            targets.stride(0), targets.stride(1), BLANK, int(FAST_MATH), nll.data_ptr(), gfac.data_ptr(),
            _stream()), "ctc2d_forward_train")
    return nll, gfac


def ctc2d_backward_apply(grad_out, log_probs, gfac):
    T, H, N, C = log_probs.shape
    grad = torch.empty_like(log_probs)
    if grad_out.dtype != log_probs.dtype:
        grad_out = grad_out.to(log_probs.dtype)
    with torch.cuda.device(log_probs.device):
        _lib.check(_lib.lib().mr_ctc2d_backward_apply_f32(
            grad_out.data_ptr(), grad_out.stride(0) if grad_out.dim() else 0, log_probs.data_ptr(),
            gfac.data_ptr(), T, H, N, C, int(FAST_MATH), grad.data_ptr(), _stream()), "ctc2d_backward_apply")
    return grad


class CTCLoss2DFunction(Function):
    """ops/ctc_2d/ctc_loss_2d.py:7-34: forward returns nll[N] (no reduction, no zero_infinity);
    backward returns (grad_log_probs, None, None, None, None)."""

    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank=0):
        ctx.blank = blank
        if not log_probs.is_cuda:
            raise NotImplementedError                                         # ctc_loss_2d.py:12-13
        ctx.fused = (not CONTRACT_PATH) and log_probs.dtype == torch.float32 and log_probs.requires_grad
        if ctx.fused:
            nll, gfac = ctc2d_forward_train(log_probs, targets, input_lengths, target_lengths, blank)
            ctx.save_for_backward(log_probs, gfac)
            return nll
        nll, log_alpha = ctc2d_forward(log_probs, targets, input_lengths, target_lengths, blank,
                                       torch.finfo().tiny)
        if log_probs.requires_grad:
            ctx.save_for_backward(log_probs, targets, input_lengths, target_lengths, nll, log_alpha)
        return nll

    @staticmethod
    def backward(ctx, grad_output):
        grad_log_probs = None
        if ctx.fused:
            log_probs, gfac = ctx.saved_tensors
            if ctx.needs_input_grad[0]:
                grad_log_probs = ctc2d_backward_apply(grad_output.contiguous(), log_probs, gfac)
        else:
            log_probs, targets, input_lengths, target_lengths, nll, log_alpha = ctx.saved_tensors
            if ctx.needs_input_grad[0]:
                grad_log_probs = ctc2d_backward(grad_output.contiguous(), log_probs, targets, input_lengths,
                                                target_lengths, nll, log_alpha, ctx.blank)
        return grad_log_probs, None, None, None, None


ctc_loss_2d = CTCLoss2DFunction.apply


class CTCLoss2D(torch.nn.Module):
    """decoders/ctc_loss2d.py:8-154 surface: forward(mask[T,H,N], classify[T,H,N,C], targets, input_lengths,
    target_lengths) with reduction none | mean (per-sample / target_length, NOT batch-averaged, :150-151) | sum.
    Computed by the CUDA op on log_probs = mask + classify (no fp32 saturation floor, unlike the python
    teaching implementation — SURVEY.md §8c)."""

    def __init__(self, blank=0, reduction='mean'):
        super().__init__()
        self.blank = blank
        self.reduction = reduction

    def forward(self, mask, classify, targets, input_lengths, target_lengths):
        log_probs = (mask.unsqueeze(-1) + classify).contiguous()
        dev = log_probs.device
        nll = ctc_loss_2d(log_probs, targets.long().to(dev), input_lengths.long().to(dev),
                          target_lengths.long().to(dev), self.blank)
        if self.reduction == 'mean':
            return nll / target_lengths.to(dev).type(nll.dtype)
        if self.reduction == 'sum':
            return nll.sum()
        return nll


CTC2DLoss = CTCLoss2D  # the name BASELINE.json uses

"""Greedy decoders on the GPU: the integer core of the reference's representers —
`CTCRepresenter.represent` (structure/representers/ctc_representer.py:22-34), `CTCRepresenter2D.represent`
(ctc_representer2d.py:27-51) and `SequenceRecognitionRepresenter.represent` (sequence_recognition_representer.py:23-28).
They return int32 label tensors; turning labels into strings stays host code (charset.label_to_string)."""
import torch

from . import _lib


def _st():
    return torch.cuda.current_stream().cuda_stream


def ctc_greedy_decode(prob, blank=0, unknown=1):
    """prob (N, C, 1, W) class scores (CRNNDecoder eval output) -> int32 (N, W) collapsed labels, blank-padded."""
    if not prob.is_cuda:
        raise NotImplementedError("megreader_b200.decode: CUDA tensors only")
    prob = prob.float()
    N, C, H, W = prob.shape
    out = torch.empty((N, W), dtype=torch.int32, device=prob.device)
    _lib.check(_lib.lib().mr_ctc_greedy_decode(prob.data_ptr(), None, N, C, 1, W, prob.stride(0), prob.stride(1),
                                               prob.stride(2), prob.stride(3), 0, 0, 0, blank, unknown, out.data_ptr(),
                                               _st()), "ctc_greedy_decode")
    return out


def ctc2d_greedy_decode(classify, mask, blank=0, unknown=1):
    """classify (N, C, H, W), mask (N, 1, H, W) (CTCDecoder2D eval output) -> int32 (N, W)."""
    if not classify.is_cuda:
        raise NotImplementedError("megreader_b200.decode: CUDA tensors only")
    classify, mask = classify.float(), mask.float()
    N, C, H, W = classify.shape
    out = torch.empty((N, W), dtype=torch.int32, device=classify.device)
    _lib.check(_lib.lib().mr_ctc_greedy_decode(classify.data_ptr(), mask.data_ptr(), N, C, H, W, classify.stride(0),
                                               classify.stride(1), classify.stride(2), classify.stride(3),
                                               mask.stride(0), mask.stride(2), mask.stride(3), blank, unknown,
                                               out.data_ptr(), _st()), "ctc2d_greedy_decode")
    return out


def blank_after_first_blank_(pred, blank=0):
    """In place on an int32 (N, W) tensor: everything from the first blank on becomes blank."""
    assert pred.dtype == torch.int32 and pred.is_contiguous() and pred.is_cuda
    _lib.check(_lib.lib().mr_blank_after_first_blank(pred.data_ptr(), pred.size(0), pred.size(1), blank, _st()),
               "blank_after_first_blank")
    return pred

"""GPU: greedy CTC decoders — bit-exact label indices against the python loops of the reference's representers
(restated in oracle/crnn_port.py::greedy_ctc_decode and here for the 2D / attention variants)."""
import numpy as np
import pytest
import torch

from oracle.crnn_port import greedy_ctc_decode

pytestmark = pytest.mark.gpu


def _collapse(pred, blank=0, unknown=1):
    out = torch.zeros(pred.shape, dtype=torch.int32) + blank
    for i in range(pred.shape[0]):
        valid, previous = 0, blank
        for j in range(pred.shape[1]):
            c = int(pred[i][j])
            if c == previous or c == unknown:
                continue
            if c != blank:
                out[i][valid] = c
                valid += 1
            previous = c
    return out


def test_ctc_greedy_decode_bit_exact(cuda):
    from megreader_b200 import decode
    torch.manual_seed(0)
    logits = torch.randn(16, 38, 1, 65) * 3
    logits[:, 0] += 2.0                      # plenty of blanks and repeats
    logits[:, 1] += 1.0                      # and unknowns
    prob = torch.softmax(logits, dim=1)
    want = greedy_ctc_decode(prob)
    got = decode.ctc_greedy_decode(prob.to(cuda)).cpu()
    assert torch.equal(got, want)
    # strided view (N, C, 1, T) of a (T, N, C) tensor, like the CRNN eval branch produces
    tnc = prob.squeeze(2).permute(2, 0, 1).contiguous().to(cuda)
    got2 = decode.ctc_greedy_decode(tnc.permute(1, 2, 0).unsqueeze(2)).cpu()
    assert torch.equal(got2, want)


def test_ctc2d_greedy_decode_bit_exact(cuda):
    from megreader_b200 import decode
    torch.manual_seed(1)
    N, C, H, W = 9, 38, 8, 32
    classify = torch.softmax(torch.randn(N, C, H, W) * 2, dim=1)
    mask = torch.softmax(torch.randn(N, 1, H, W) * 2, dim=2)
    heatmap = classify * mask                                     # ctc_representer2d.py:27-35
    paths = heatmap.max(1, keepdim=True)[0].argmax(2, keepdim=True).repeat(1, C, 1, 1)
    pred = heatmap.gather(2, paths).argmax(1).squeeze(1)
    want = _collapse(pred)
    got = decode.ctc2d_greedy_decode(classify.to(cuda), mask.to(cuda)).cpu()
    assert torch.equal(got, want)


def test_blank_after_first_blank(cuda):
    from megreader_b200 import decode
    rng = np.random.RandomState(2)
    pred = torch.from_numpy(rng.randint(0, 5, size=(33, 32)).astype(np.int32))
    want = pred.clone()
    m = torch.ones(33, dtype=torch.int32)
    for i in range(32):                                           # sequence_recognition_representer.py:23-28
        m = (1 - (want[:, i] == 0).int()) * m
        want[:, i] = want[:, i] * m
    got = decode.blank_after_first_blank_(pred.to(cuda).contiguous()).cpu()
    assert torch.equal(got, want)

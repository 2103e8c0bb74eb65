"""GPU: the batched input step (csrc/input_pipeline.cu through megreader_b200.input_pipeline) against the goldens recorded from
the UNMODIFIED reference processes and against the CPU oracle.  The per-pixel arithmetic itself is already checked on the CPU
(tests/test_input_core_host.py compiles the same routine for the host); this file covers the launch glue and the host API.
(Named to sort last: added after the round's GPU budget was spent, so it is the one GPU test not yet run on hardware.)"""
import os

import numpy as np
import pytest
import torch

from megreader_b200 import input_pipeline
from megreader_b200.charset import EnglishCharset
from oracle import input_port
from tests.input_cases import MODES, TOL, input_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "input_ref.npz")


@pytest.mark.parametrize("mode", sorted(MODES))
@pytest.mark.parametrize("as_u8", [True, False])
def test_resize_normalize_batch(cuda, mode, as_u8):
    images, _ = input_cases()
    batch = images if as_u8 else [im.astype(np.float32) for im in images]
    out = input_pipeline.resize_normalize(batch, MODES[mode], mode, device=cuda)
    assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (len(images), 3) + MODES[mode]
    got = out.cpu().numpy()
    np.testing.assert_allclose(got, np.load(GOLD)["image." + mode], rtol=0, atol=TOL)
    want = np.stack([input_port.normalize(input_port.resize_or_pad(im.astype(np.float32), MODES[mode], mode)) for im in images])
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)              # device FMA contraction vs separate roundings


def test_pack_labels_batch(cuda):
    _, texts = input_cases()
    labels, lengths = input_pipeline.pack_labels(texts, EnglishCharset(), 32, device=cuda)
    g = np.load(GOLD)
    assert labels.dtype == torch.int32 and np.array_equal(labels.cpu().numpy(), g["labels"])
    assert np.array_equal(lengths.cpu().numpy(), g["lengths"])


def test_empty_batches_and_cpu_refusal(cuda):
    assert tuple(input_pipeline.resize_normalize([], (32, 128), device=cuda).shape) == (0, 3, 32, 128)
    labels, lengths = input_pipeline.pack_labels([], device=cuda)
    assert tuple(labels.shape) == (0, 32) and tuple(lengths.shape) == (0,)
    with pytest.raises(NotImplementedError):
        input_pipeline.resize_normalize([np.zeros((4, 4, 3), np.uint8)], (32, 128), device="cpu")

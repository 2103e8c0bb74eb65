"""CPU: the C oracle of the 2D-CTC op against the golden vectors generated from the reference's own
python CTCLoss2D (oracle/make_golden.py), plus self-consistency of the K3 restatement."""
import glob
import os

import numpy as np
import pytest

from oracle import capi
from tests.cases import ctc2d_case

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ctc2d_pyref_*.npz")))


def test_golden_present():
    assert len(GOLD) >= 5


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_oracle_matches_reference_python_nll(path, dtype):
    d = np.load(path)
    nll, _ = capi.ctc2d_forward(d["log_probs"].astype(dtype), d["targets"], d["input_lengths"], d["target_lengths"])
    np.testing.assert_allclose(nll, d["ref_nll"], rtol=2e-6 if dtype == np.float64 else 2e-5)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_grad_is_reference_autograd_plus_exp_lp(path):
    """K3 returns exp(lp) - exp(G + nll - lp) on in-target classes, 0 elsewhere;
    the reference python loss's autograd gives -exp(G + nll - lp)."""
    d = np.load(path)
    lp = d["log_probs"].astype(np.float64)
    nll, la = capi.ctc2d_forward(lp, d["targets"], d["input_lengths"], d["target_lengths"])
    g = capi.ctc2d_backward(np.ones_like(nll), lp, d["targets"], d["input_lengths"], d["target_lengths"], nll, la)
    in_target = np.zeros(lp.shape, bool)
    for b in range(lp.shape[2]):
        cls = set(d["targets"][b, :d["target_lengths"][b]].tolist()) | {0}
        in_target[:, :, b, sorted(cls)] = True
    assert np.all(g[~in_target] == 0)
    # in-target classes whose alpha+beta mass is -inf at (t,h) also get exactly 0 (kernel.cu:506-507)
    true_grad = np.where(g != 0, g - np.exp(lp), 0.0)
    np.testing.assert_allclose(true_grad, d["ref_autograd"], atol=5e-6)


def test_oracle_f32_vs_f64():
    lp, tg, il, tl = ctc2d_case(7, 16, 4, 5, 11, 8, ragged_T=True)
    n64, a64 = capi.ctc2d_forward(lp.astype(np.float64), tg, il, tl)
    n32, a32 = capi.ctc2d_forward(lp, tg, il, tl)
    np.testing.assert_allclose(n32, n64, rtol=1e-5)
    fin = np.isfinite(a64)
    assert np.array_equal(fin, np.isfinite(a32))
    np.testing.assert_allclose(a32[fin], a64[fin], rtol=1e-4, atol=1e-4)


def test_oracle_h1_equals_torch_ctc():
    """H = 1 degenerates to 1D CTC: compare with torch's CPU ctc_loss (third-party arithmetic the reference's
    CRNN head uses, decoders/crnn.py:47-48)."""
    import torch
    lp, tg, il, tl = ctc2d_case(11, 20, 1, 6, 9, 7, ragged_T=True, dtype=np.float64)
    nll, _ = capi.ctc2d_forward(lp, tg, il, tl)
    ref = torch.nn.functional.ctc_loss(torch.from_numpy(lp[:, 0]), torch.from_numpy(tg), torch.from_numpy(il),
                                       torch.from_numpy(tl), blank=0, reduction="none")
    np.testing.assert_allclose(nll, ref.numpy(), rtol=1e-9)

"""CPU: the C oracle of DCNv1/v2 pinned against torchvision.ops.deform_conv2d (third-party, same mmdetection
lineage as assets/ops/dcn; valid when offset spatial size == output size) — forward and all
backward products through torch autograd."""
import numpy as np
import pytest
import torch

from oracle import capi

tv = pytest.importorskip("torchvision.ops")

CASES = [
    # B, C, H, W, Cout, k, stride, pad, dil, group, dg, modulated, bias
    (2, 4, 7, 9, 6, 3, 1, 1, 1, 1, 1, True, False),
    (2, 4, 8, 8, 4, 3, 2, 1, 1, 1, 1, True, True),
    (1, 8, 6, 5, 8, 3, 1, 1, 1, 2, 2, True, True),
    (2, 6, 9, 7, 4, 3, 1, 2, 2, 1, 3, True, False),
    (2, 4, 7, 9, 6, 3, 1, 1, 1, 1, 1, False, False),
    (3, 4, 5, 5, 2, 1, 1, 0, 1, 1, 1, True, True),
]


def _inputs(seed, B, C, H, W, Cout, k, s, p, d, group, dg):
    rng = np.random.RandomState(seed)
    Ho = (H + 2 * p - (d * (k - 1) + 1)) // s + 1
    Wo = (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    x = rng.standard_normal((B, C, H, W))
    w = rng.standard_normal((Cout, C // group, k, k)) * 0.3
    b = rng.standard_normal((Cout,))
    off = rng.standard_normal((B, 2 * k * k * dg, Ho, Wo)) * 1.5
    m = 1 / (1 + np.exp(-rng.standard_normal((B, k * k * dg, Ho, Wo))))
    go = rng.standard_normal((B, Cout, Ho, Wo))
    return x, w, b, off, m, go


@pytest.mark.parametrize("case", CASES)
def test_dcn_oracle_vs_torchvision(case):
    B, C, H, W, Cout, k, s, p, d, group, dg, modulated, with_bias = case
    x, w, b, off, m, go = _inputs(0, B, C, H, W, Cout, k, s, p, d, group, dg)
    tx, tw, tb, toff, tm = [torch.tensor(a, requires_grad=True) for a in (x, w, b, off, m)]
    out = tv.deform_conv2d(tx, toff, tw, tb if with_bias else None, stride=s, padding=p, dilation=d,
                           mask=tm if modulated else None)
    out.backward(torch.tensor(go))
    o = capi.dcn_forward(x, w, b if with_bias else None, off, m if modulated else None, s, p, d, group, dg)
    np.testing.assert_allclose(o, out.detach().numpy(), rtol=1e-9, atol=1e-10)
    gi, gw, gb, goff, gm = capi.dcn_backward(x, w, b if with_bias else None, off, m if modulated else None, go,
                                             s, p, d, group, dg)
    np.testing.assert_allclose(gi, tx.grad.numpy(), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(gw, tw.grad.numpy(), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(goff, toff.grad.numpy(), rtol=1e-8, atol=1e-9)
    if modulated:
        np.testing.assert_allclose(gm, tm.grad.numpy(), rtol=1e-8, atol=1e-9)
    if with_bias:
        np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-8, atol=1e-9)


def test_dcn_oracle_flat_offset_reindexing():
    """Offset map larger than the output (stride-2 Bottleneck, backbones/resnet.py:136-142): each sample's slab is
    read flat with (Ho,Wo) strides (deform_conv_cuda_kernel.cu:599-609) == using the first 18*Ho*Wo values."""
    B, C, H, W, Cout, k, s, p = 2, 4, 8, 8, 4, 3, 2, 1
    rng = np.random.RandomState(3)
    x = rng.standard_normal((B, C, H, W))
    w = rng.standard_normal((Cout, C, k, k))
    off_big = rng.standard_normal((B, 18, H, W))
    m_big = rng.uniform(size=(B, 9, H, W))
    Ho = Wo = 4
    off_small = off_big.reshape(B, -1)[:, :18 * Ho * Wo].reshape(B, 18, Ho, Wo)
    m_small = m_big.reshape(B, -1)[:, :9 * Ho * Wo].reshape(B, 9, Ho, Wo)
    a = capi.dcn_forward(x, w, None, off_big, m_big, s, p, 1)
    b = capi.dcn_forward(x, w, None, off_small, m_small, s, p, 1)
    assert np.array_equal(a, b)
    go = rng.standard_normal(a.shape)
    ga = capi.dcn_backward(x, w, None, off_big, m_big, go, s, p, 1)
    gb = capi.dcn_backward(x, w, None, off_small, m_small, go, s, p, 1)
    assert np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1])
    assert np.array_equal(ga[3].reshape(B, -1)[:, :18 * Ho * Wo], gb[3].reshape(B, -1))
    assert np.all(ga[3].reshape(B, -1)[:, 18 * Ho * Wo:] == 0)   # tail of each slab stays zero (App. B2.1)

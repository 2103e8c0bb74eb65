"""CPU: the deformable PS-RoI pooling oracle (oracle/deform_pool_oracle.c).  The reference ships no test, golden or CPU
path for this op (PARITY UNPINNED, see the oracle header); what can be pinned is internal: the backward is the derivative of
the forward (central finite differences in fp64), and closed-form cases."""
import numpy as np
import pytest

from oracle import capi
from tests.deform_pool_cases import CASES, make


def test_constant_map_pools_to_the_constant():
    data, rois, trans, args = make("trans")
    data[:] = 2.5
    out, cnt = capi.deform_psroi_forward(data, rois, trans, **args)
    assert np.all((cnt > 0) | (out == 0))
    np.testing.assert_allclose(out[cnt > 0], 2.5, rtol=1e-12)


def test_outside_roi_counts_zero_and_gets_no_gradient():
    data, rois, trans, args = make("outside")
    out, cnt = capi.deform_psroi_forward(data, rois, trans, **args)
    assert np.all(cnt[0] == 0) and np.all(out[0] == 0)
    go = np.ones_like(out)
    go[1:] = 0
    gin, gtr = capi.deform_psroi_backward(go, data, rois, trans, cnt, **args)
    assert not gin.any() and not gtr.any()


@pytest.mark.parametrize("name", sorted(CASES))
def test_backward_is_the_derivative_of_forward(name):
    data, rois, trans, args = make(name)
    rng = np.random.RandomState(5)
    out, cnt = capi.deform_psroi_forward(data, rois, trans, **args)
    go = rng.standard_normal(out.shape)
    gin, gtr = capi.deform_psroi_backward(go, data, rois, trans, cnt, **args)

    def loss(d, t):
        o, _ = capi.deform_psroi_forward(d, rois, t, **args)
        return float((o * go).sum())
    eps = 1e-6
    for _ in range(12):                                    # data gradient (the op is linear in data: exact up to rounding)
        idx = tuple(rng.randint(s) for s in data.shape)
        d1, d2 = data.copy(), data.copy()
        d1[idx] += eps
        d2[idx] -= eps
        np.testing.assert_allclose((loss(d1, trans) - loss(d2, trans)) / (2 * eps), gin[idx], rtol=1e-6, atol=1e-8)
    if trans is not None and args["trans_std"] > 0:
        checked = 0
        for _ in range(40):                                # offset gradient: piecewise smooth, skip kinks
            idx = tuple(rng.randint(s) for s in trans.shape)
            t1, t2 = trans.copy(), trans.copy()
            t1[idx] += eps
            t2[idx] -= eps
            c1 = capi.deform_psroi_forward(data, rois, t1, **args)[1]
            c2 = capi.deform_psroi_forward(data, rois, t2, **args)[1]
            if not (np.array_equal(c1, cnt) and np.array_equal(c2, cnt)):
                continue                                   # a sample crossed the border inside the step
            fd = (loss(data, t1) - loss(data, t2)) / (2 * eps)
            if abs(fd - gtr[idx]) > 1e-4 * max(1.0, abs(fd)):
                # a sample sits within eps of a pixel boundary (bilinear kink): retry with a smaller step must agree
                t1[idx] = trans[idx] + eps * 1e-2
                t2[idx] = trans[idx] - eps * 1e-2
                fd = (loss(data, t1) - loss(data, t2)) / (2 * eps * 1e-2)
            np.testing.assert_allclose(fd, gtr[idx], rtol=2e-3, atol=1e-5)
            checked += 1
        assert checked >= 10


def test_float_build_matches_double_build():
    data, rois, trans, args = make("parts")
    o64, c64 = capi.deform_psroi_forward(data, rois, trans, **args)
    o32, c32 = capi.deform_psroi_forward(data.astype(np.float32), rois.astype(np.float32), trans.astype(np.float32), **args)
    assert np.mean(c32 != c64) < 0.02                       # a sample exactly on the border may flip in fp32
    same = c32 == c64
    np.testing.assert_allclose(o32[same], o64[same], rtol=2e-4, atol=2e-5)

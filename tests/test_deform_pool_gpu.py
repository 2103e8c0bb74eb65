"""GPU: deformable PS-RoI pooling (csrc/deform_pool.cu through the C-ABI and the reference-named surfaces) against the CPU
oracle (oracle/deform_pool_oracle.c, fp64) on the shared seeded cases, plus the Pack modules' zero-initialised branches."""
import numpy as np
import pytest
import torch

from oracle import capi
from tests.deform_pool_cases import CASES, make

pytestmark = pytest.mark.gpu


def _api():
    from megreader_b200 import refapi
    refapi.install()
    import assets.ops.dcn as dcn
    assert "refapi" in dcn.__file__
    return dcn


@pytest.mark.parametrize("name", sorted(CASES))
def test_forward_backward_vs_oracle(cuda, name):
    dcn = _api()
    data, rois, trans, a = make(name)
    out_ref, cnt_ref = capi.deform_psroi_forward(data, rois, trans, **a)
    rng = np.random.RandomState(1)
    go = rng.standard_normal(out_ref.shape)
    gin_ref, gtr_ref = capi.deform_psroi_backward(go, data, rois, trans, cnt_ref, **a)
    d = torch.from_numpy(data).float().to(cuda).requires_grad_(True)
    r = torch.from_numpy(rois).float().to(cuda)
    t = (torch.from_numpy(trans).float().to(cuda).requires_grad_(True) if trans is not None else d.new_empty(0))
    out = dcn.deform_roi_pooling(d, r, t, a["spatial_scale"], a["pooled"], a["output_dim"], a["no_trans"], a["group_size"],
                                 a["part_size"], a["sample_per_part"], a["trans_std"])
    assert tuple(out.shape) == out_ref.shape
    # a sample that sits on the half-pixel border within fp32 rounding may be counted on one side only: such bins (and
    # the gradients they touch) are excluded; they must be rare
    out.backward(torch.from_numpy(go).float().to(cuda))
    got = out.detach().cpu().numpy()
    bad = np.abs(got - out_ref) > 1e-4 * (1 + np.abs(out_ref))
    assert bad.mean() < 0.02
    if not bad.any():
        np.testing.assert_allclose(d.grad.cpu().numpy(), gin_ref, rtol=1e-3, atol=2e-4)
        if trans is not None:
            np.testing.assert_allclose(t.grad.cpu().numpy(), gtr_ref, rtol=2e-3, atol=2e-3 * max(1.0, np.abs(gtr_ref).max()))


def test_empty_rois_and_cpu_refusal(cuda):
    dcn = _api()
    d = torch.randn(1, 4, 8, 8, device=cuda)
    out = dcn.deform_roi_pooling(d, d.new_zeros(0, 5), d.new_empty(0), 1.0, 2, 4, True)
    assert tuple(out.shape) == (0, 4, 2, 2)
    with pytest.raises(NotImplementedError):
        dcn.deform_roi_pooling(d.cpu(), torch.zeros(1, 5), torch.empty(0), 1.0, 2, 4, True)


def test_pack_modules_zero_initialised_branches(cuda):
    """DeformRoIPoolingPack / ModulatedDeformRoIPoolingPack start with zero offsets (and mask = sigmoid(0) = 1/2):
    modules/deform_pool.py:66-68,146-148."""
    dcn = _api()
    torch.manual_seed(0)
    d = torch.randn(2, 4, 12, 12, device=cuda)
    rois = torch.tensor([[0, 1.2, 2.1, 8.3, 9.4], [1, 0.4, 0.2, 5.5, 6.6], [0, 3.3, 1.1, 10.2, 7.7]], device=cuda)
    plain = dcn.DeformRoIPooling(1.0, 3, 4, True)(d, rois, None)
    pack = dcn.DeformRoIPoolingPack(1.0, 3, 4, False, trans_std=0.1, deform_fc_channels=16).to(cuda)
    torch.testing.assert_close(pack(d, rois), plain)
    mod = dcn.ModulatedDeformRoIPoolingPack(1.0, 3, 4, False, trans_std=0.1, deform_fc_channels=16).to(cuda)
    torch.testing.assert_close(mod(d, rois), 0.5 * plain)
    assert sorted(k.split(".")[0] for k in mod.state_dict()) == ["mask_fc"] * 4 + ["offset_fc"] * 6
    dd = d.clone().requires_grad_(True)
    mod(dd, rois).square().sum().backward()
    assert torch.isfinite(dd.grad).all() and all(p.grad is not None for p in mod.parameters())

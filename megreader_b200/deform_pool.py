"""Deformable position-sensitive RoI pooling with the reference's surface:
  extension functions  deform_psroi_pooling_cuda_forward / _backward   (assets/ops/dcn/src/deform_pool_cuda.cpp:29-81)
  DeformRoIPoolingFunction / deform_roi_pooling                         (functions/deform_pool.py:7-69)
  DeformRoIPooling, DeformRoIPoolingPack, ModulatedDeformRoIPoolingPack (modules/deform_pool.py:6-172)
CUDA only (csrc/deform_pool.cu through the C-ABI), fp32; no CPU fallback."""
import torch
from torch import nn
from torch.autograd import Function

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(data, what):
    if not data.is_cuda:
        raise NotImplementedError
    if data.dtype != torch.float32:
        raise RuntimeError("megreader_b200: %s is built for float32" % what)


def deform_psroi_pooling_cuda_forward(input, bbox, trans, out, top_count, no_trans, spatial_scale, output_dim, group_size,
                                      pooled_size, part_size, sample_per_part, trans_std):
    _check(input, "deform_psroi_pooling_cuda_forward")
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")                       # deform_pool_cuda.cpp:34
    batch, channels, height, width = input.shape
    channels_trans = 2 if no_trans else trans.size(1)
    num_bbox = bbox.size(0)
    if num_bbox != out.size(0):
        raise RuntimeError("Output shape and bbox number wont match: (%d vs %d)." % (out.size(0), num_bbox))
    bbox = bbox.contiguous().float()
    trans_c = None if no_trans else trans.contiguous()
    with torch.cuda.device(input.device):
        _lib.check(_lib.lib().mr_deform_psroi_pool_forward_f32(
            input.data_ptr(), bbox.data_ptr(), trans_c.data_ptr() if trans_c is not None else None, batch, channels,
            height, width, num_bbox, channels_trans, int(bool(no_trans)), float(spatial_scale), output_dim, group_size,
            pooled_size, part_size, sample_per_part, float(trans_std), out.data_ptr(), top_count.data_ptr(), _stream()),
            "deform_psroi_pool_forward")


def deform_psroi_pooling_cuda_backward(out_grad, input, bbox, trans, top_count, input_grad, trans_grad, no_trans,
                                       spatial_scale, output_dim, group_size, pooled_size, part_size, sample_per_part,
                                       trans_std):
    _check(out_grad, "deform_psroi_pooling_cuda_backward")
    if not out_grad.is_contiguous():
        raise RuntimeError("out_grad tensor has to be contiguous")                    # deform_pool_cuda.cpp:59
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")
    batch, channels, height, width = input.shape
    channels_trans = 2 if no_trans else trans.size(1)
    num_bbox = bbox.size(0)
    if num_bbox != out_grad.size(0):
        raise RuntimeError("Output shape and bbox number wont match: (%d vs %d)." % (out_grad.size(0), num_bbox))
    bbox = bbox.contiguous().float()
    trans_c = None if no_trans else trans.contiguous()
    with torch.cuda.device(input.device):
        _lib.check(_lib.lib().mr_deform_psroi_pool_backward_f32(
            out_grad.data_ptr(), input.data_ptr(), bbox.data_ptr(), trans_c.data_ptr() if trans_c is not None else None,
            top_count.data_ptr(), batch, channels, height, width, num_bbox, channels_trans, int(bool(no_trans)),
            float(spatial_scale), output_dim, group_size, pooled_size, part_size, sample_per_part, float(trans_std),
            input_grad.data_ptr(), None if no_trans else trans_grad.data_ptr(), _stream()), "deform_psroi_pool_backward")


class DeformRoIPoolingFunction(Function):
    @staticmethod
    def forward(ctx, data, rois, offset, spatial_scale, out_size, out_channels, no_trans, group_size=1, part_size=None,
                sample_per_part=4, trans_std=.0):
        ctx.spatial_scale, ctx.out_size, ctx.out_channels, ctx.no_trans = spatial_scale, out_size, out_channels, no_trans
        ctx.group_size = group_size
        ctx.part_size = out_size if part_size is None else part_size
        ctx.sample_per_part, ctx.trans_std = sample_per_part, trans_std
        assert 0.0 <= ctx.trans_std <= 1.0
        if not data.is_cuda:
            raise NotImplementedError
        n = rois.shape[0]
        output = data.new_empty(n, out_channels, out_size, out_size)
        output_count = data.new_empty(n, out_channels, out_size, out_size)
        deform_psroi_pooling_cuda_forward(data, rois, offset, output, output_count, ctx.no_trans, ctx.spatial_scale,
                                          ctx.out_channels, ctx.group_size, ctx.out_size, ctx.part_size,
                                          ctx.sample_per_part, ctx.trans_std)
        if data.requires_grad or rois.requires_grad or offset.requires_grad:
            ctx.save_for_backward(data, rois, offset)
        ctx.output_count = output_count
        return output

    @staticmethod
    def backward(ctx, grad_output):
        if not grad_output.is_cuda:
            raise NotImplementedError
        data, rois, offset = ctx.saved_tensors
        grad_input = torch.zeros_like(data)
        grad_offset = torch.zeros_like(offset)
        deform_psroi_pooling_cuda_backward(grad_output.contiguous(), data, rois, offset, ctx.output_count, grad_input,
                                           grad_offset, ctx.no_trans, ctx.spatial_scale, ctx.out_channels, ctx.group_size,
                                           ctx.out_size, ctx.part_size, ctx.sample_per_part, ctx.trans_std)
        return (grad_input, None, grad_offset, None, None, None, None, None, None, None, None)


deform_roi_pooling = DeformRoIPoolingFunction.apply


class DeformRoIPooling(nn.Module):
    def __init__(self, spatial_scale, out_size, out_channels, no_trans, group_size=1, part_size=None, sample_per_part=4,
                 trans_std=.0):
        super().__init__()
        self.spatial_scale, self.out_size, self.out_channels, self.no_trans = spatial_scale, out_size, out_channels, no_trans
        self.group_size = group_size
        self.part_size = out_size if part_size is None else part_size
        self.sample_per_part, self.trans_std = sample_per_part, trans_std

    def _pool(self, data, rois, offset, no_trans):
        return deform_roi_pooling(data, rois, offset, self.spatial_scale, self.out_size, self.out_channels, no_trans,
                                  self.group_size, self.part_size, self.sample_per_part, self.trans_std)

    def forward(self, data, rois, offset):
        if self.no_trans:
            offset = data.new_empty(0)
        return self._pool(data, rois, offset, self.no_trans)


def _fc_stack(in_features, hidden, out_features, count, final=None):
    """count Linear layers: hidden width in between, ReLU after all but the last, optional module after the last; the last
    Linear is zero-initialised (modules/deform_pool.py:56-68,120-148)."""
    seq, ic = [], in_features
    for i in range(count):
        oc = hidden if i < count - 1 else out_features
        seq.append(nn.Linear(ic, oc))
        ic = oc
        if i < count - 1:
            seq.append(nn.ReLU(inplace=True))
    last = seq[-1]
    last.weight.data.zero_()
    last.bias.data.zero_()
    if final is not None:
        seq.append(final)
    return nn.Sequential(*seq)


class DeformRoIPoolingPack(DeformRoIPooling):
    def __init__(self, spatial_scale, out_size, out_channels, no_trans, group_size=1, part_size=None, sample_per_part=4,
                 trans_std=.0, num_offset_fcs=3, deform_fc_channels=1024):
        super().__init__(spatial_scale, out_size, out_channels, no_trans, group_size, part_size, sample_per_part, trans_std)
        self.num_offset_fcs, self.deform_fc_channels = num_offset_fcs, deform_fc_channels
        if not no_trans:
            cells = self.out_size * self.out_size
            self.offset_fc = _fc_stack(cells * self.out_channels, deform_fc_channels, cells * 2, num_offset_fcs)

    def _offsets(self, data, rois):
        n = rois.shape[0]
        pooled = self._pool(data, rois, data.new_empty(0), True)                  # plain PS-RoI pooling feeds the FCs
        return pooled, self.offset_fc(pooled.view(n, -1)).view(n, 2, self.out_size, self.out_size)

    def forward(self, data, rois):
        assert data.size(1) == self.out_channels
        if self.no_trans:
            return self._pool(data, rois, data.new_empty(0), True)
        _, offset = self._offsets(data, rois)
        return self._pool(data, rois, offset, False)


class ModulatedDeformRoIPoolingPack(DeformRoIPoolingPack):
    def __init__(self, spatial_scale, out_size, out_channels, no_trans, group_size=1, part_size=None, sample_per_part=4,
                 trans_std=.0, num_offset_fcs=3, num_mask_fcs=2, deform_fc_channels=1024):
        super().__init__(spatial_scale, out_size, out_channels, no_trans, group_size, part_size, sample_per_part, trans_std,
                         num_offset_fcs, deform_fc_channels)
        self.num_mask_fcs = num_mask_fcs
        if not no_trans:
            cells = self.out_size * self.out_size
            self.mask_fc = _fc_stack(cells * self.out_channels, deform_fc_channels, cells, num_mask_fcs, nn.Sigmoid())

    def forward(self, data, rois):
        assert data.size(1) == self.out_channels
        if self.no_trans:
            return self._pool(data, rois, data.new_empty(0), True)
        n = rois.shape[0]
        pooled, offset = self._offsets(data, rois)
        mask = self.mask_fc(pooled.view(n, -1)).view(n, 1, self.out_size, self.out_size)
        return self._pool(data, rois, offset, False) * mask

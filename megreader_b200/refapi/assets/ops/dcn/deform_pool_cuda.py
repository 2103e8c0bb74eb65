"""Stands where the reference's extension `deform_pool_cuda*.so` stands (assets/ops/dcn/setup.py:11-14); same two exports
as assets/ops/dcn/src/deform_pool_cuda.cpp:83-86."""
from megreader_b200.deform_pool import deform_psroi_pooling_cuda_backward, deform_psroi_pooling_cuda_forward  # noqa: F401

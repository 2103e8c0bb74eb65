"""Stands where the reference's extension `deform_conv_cuda*.so` stands (assets/ops/dcn/setup.py:7-10);
same five exports as assets/ops/dcn/src/deform_conv_cuda.cpp:681-695."""
from megreader_b200.dcn import (deform_conv_backward_input_cuda, deform_conv_backward_parameters_cuda,  # noqa: F401
                                deform_conv_forward_cuda, modulated_deform_conv_cuda_backward,
                                modulated_deform_conv_cuda_forward)

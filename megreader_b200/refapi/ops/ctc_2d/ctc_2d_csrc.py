"""Stands where the reference's pybind extension `ctc_2d_csrc*.so` stands
(ops/ctc_2d/setup.py:16,47; exports ctc2d_forward / ctc2d_backward, ops/ctc_2d/csrc/ctc2d.cpp:3-6)."""
from megreader_b200.ctc2d import ctc2d_backward, ctc2d_forward  # noqa: F401

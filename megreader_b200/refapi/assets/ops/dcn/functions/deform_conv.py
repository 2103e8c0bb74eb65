# mirrors assets/ops/dcn/functions/deform_conv.py
from megreader_b200.dcn import (DeformConvFunction, ModulatedDeformConvFunction, deform_conv,  # noqa: F401
                                modulated_deform_conv)
from .. import deform_conv_cuda  # noqa: F401

# mirrors assets/ops/dcn/modules/deform_conv.py
from megreader_b200.dcn import DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack  # noqa: F401

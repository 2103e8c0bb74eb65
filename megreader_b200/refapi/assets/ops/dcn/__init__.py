# mirrors assets/ops/dcn/__init__.py (deform_pool is not built yet: SURVEY.md §8f row N4)
from .functions.deform_conv import deform_conv, modulated_deform_conv
from .modules.deform_conv import (DeformConv, ModulatedDeformConv,
                                  DeformConvPack, ModulatedDeformConvPack)

__all__ = ['DeformConv', 'DeformConvPack', 'ModulatedDeformConv', 'ModulatedDeformConvPack', 'deform_conv',
           'modulated_deform_conv']

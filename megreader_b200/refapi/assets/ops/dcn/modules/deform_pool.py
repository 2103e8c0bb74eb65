# mirrors assets/ops/dcn/modules/deform_pool.py
from megreader_b200.deform_pool import DeformRoIPooling, DeformRoIPoolingPack, ModulatedDeformRoIPoolingPack  # noqa: F401

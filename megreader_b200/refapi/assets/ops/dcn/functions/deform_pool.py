# mirrors assets/ops/dcn/functions/deform_pool.py
from megreader_b200.deform_pool import DeformRoIPoolingFunction, deform_roi_pooling  # noqa: F401
from .. import deform_pool_cuda  # noqa: F401

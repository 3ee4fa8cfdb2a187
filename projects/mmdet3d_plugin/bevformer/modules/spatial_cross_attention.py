"""Drop-in module path: the implementation lives in bevformer_b200/plugin/spatial_cross_attention.py."""
from bevformer_b200.plugin.spatial_cross_attention import (  # noqa: F401
    MSDeformableAttention3D, ScaPlan, SpatialCrossAttention)
from .multi_scale_deformable_attn_function import (  # noqa: F401
    MultiScaleDeformableAttnFunction_fp16, MultiScaleDeformableAttnFunction_fp32)

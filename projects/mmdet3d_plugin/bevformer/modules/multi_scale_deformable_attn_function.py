"""Drop-in module path of the op wrapper: the implementation lives in bevformer_b200/ops.py and
calls libbevformer_b200.so (bevf_msda_forward / bevf_msda_backward)."""
from bevformer_b200.ops import (  # noqa: F401
    MultiScaleDeformableAttnFunction_fp16, MultiScaleDeformableAttnFunction_fp32)

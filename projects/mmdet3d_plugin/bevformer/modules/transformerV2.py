"""Drop-in module path for BEVFormerV2's BEV-encoder wrapper (implementation:
bevformer_b200/plugin/transformer.py).  ResNetFusion / PerceptionTransformerV2's decoder half are out of scope."""
from bevformer_b200.plugin.transformer import PerceptionTransformerBEVEncoder  # noqa: F401

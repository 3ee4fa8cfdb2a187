"""Drop-in module path: the implementation lives in bevformer_b200/plugin/temporal_self_attention.py."""
from bevformer_b200.plugin.temporal_self_attention import TemporalSelfAttention  # noqa: F401

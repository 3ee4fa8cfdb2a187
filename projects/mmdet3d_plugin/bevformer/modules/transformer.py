"""Drop-in module path: the implementation lives in bevformer_b200/plugin/transformer.py
(get_bev_features only; the decoder half of the reference class is out of scope)."""
from bevformer_b200.plugin.transformer import PerceptionTransformer  # noqa: F401

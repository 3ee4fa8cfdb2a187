"""Drop-in module path: the implementation lives in bevformer_b200/plugin/encoder.py."""
from bevformer_b200.plugin.encoder import FFN, MyCustomBaseTransformerLayer  # noqa: F401

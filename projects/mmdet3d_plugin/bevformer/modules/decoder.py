"""Drop-in module path for the decoder's cross-attention (the implementation lives in
bevformer_b200/plugin/decoder.py).  DetectionTransformerDecoder itself is out of scope: a real checkout
keeps its own decoder.py and imports this class over its CustomMSDeformableAttention."""
from bevformer_b200.plugin.decoder import CustomMSDeformableAttention  # noqa: F401

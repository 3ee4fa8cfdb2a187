# Same import surface as the reference package of this path, for the encoder hot path only.
from .spatial_cross_attention import SpatialCrossAttention, MSDeformableAttention3D
from .temporal_self_attention import TemporalSelfAttention
from .encoder import BEVFormerEncoder, BEVFormerLayer
from .custom_base_transformer_layer import MyCustomBaseTransformerLayer
from .transformer import PerceptionTransformer

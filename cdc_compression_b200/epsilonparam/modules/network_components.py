"""Reference names for the building blocks (epsilonparam/modules/network_components.py).  The eps
variant spells the 7x7 switch ``large``; the shared implementation calls it ``large_filter``."""
from cdc_compression_b200._shared import layers as _L
from cdc_compression_b200._shared.layers import (  # noqa: F401
    GDN1, Downsample, FlexiblePrior, LayerNorm, LinearAttention, PreNorm, PriorFunction, Residual, Upsample,
    VBRCondition)


class Block(_L.Block):
    def __init__(self, dim, dim_out, large=False):
        super().__init__(dim, dim_out, large_filter=large)


class ResnetBlock(_L.ResnetBlock):
    def __init__(self, dim, dim_out, time_emb_dim=None, large=False):
        super().__init__(dim, dim_out, time_emb_dim, large_filter=large)

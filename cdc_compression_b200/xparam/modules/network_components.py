"""Reference names for the building blocks (xparam/modules/network_components.py)."""
from cdc_compression_b200._shared.layers import (  # noqa: F401
    GDN1, Block, Downsample, FlexiblePrior, LayerNorm, LinearAttention, PreNorm, PriorFunction, Residual,
    ResnetBlock, Upsample, VBRCondition)

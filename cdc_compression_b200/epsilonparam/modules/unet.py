"""``Unet`` of the epsilon-parameterised variant (reference epsilonparam/modules/unet.py:17-124)."""
from cdc_compression_b200._shared.unet_impl import UnetBase


class Unet(UnetBase):
    variant = "eps"

    def __init__(self, dim, out_dim=None, dim_mults=(1, 2, 4, 8), context_dim_mults=(1, 2, 3, 3), channels=3,
                 context_channels=3, with_time_emb=True):
        super().__init__()
        self._build(dim, out_dim, dim_mults, context_dim_mults, channels, context_channels, with_time_emb)

"""Context networks of the x variant (reference xparam/modules/compress_modules.py)."""
from torch import nn

from cdc_compression_b200._shared.compressor_impl import HyperpriorCompressor
from .network_components import Downsample, ResnetBlock, Upsample


class Compressor(HyperpriorCompressor):
    has_vbr_slot = False

    def __init__(self, dim=64, dim_mults=(1, 2, 3, 4), reverse_dim_mults=(4, 3, 2, 1), hyper_dims_mults=(4, 4, 4),
                 channels=3, out_channels=3):
        super().__init__()
        self._init_dims(dim, dim_mults, [dim * m for m in reverse_dim_mults] + [out_channels], hyper_dims_mults,
                        channels, out_channels)
        assert self.dims[-1] == self.reversed_dims[0]


class ResnetCompressor(Compressor):
    def __init__(self, dim=64, dim_mults=(1, 2, 3, 4), reverse_dim_mults=(4, 3, 2, 1), hyper_dims_mults=(4, 4, 4),
                 channels=3, out_channels=3):
        super().__init__(dim, dim_mults, reverse_dim_mults, hyper_dims_mults, channels, out_channels)
        self.build_network()

    def build_network(self):
        super().build_network()
        for i, (a, b) in enumerate(self.in_out):
            self.enc.append(nn.ModuleList([ResnetBlock(a, b, None, i == 0), Downsample(b)]))
        n = len(self.reversed_in_out)
        for i, (a, b) in enumerate(self.reversed_in_out):
            mid = a if i >= n - 1 else b
            self.dec.append(nn.ModuleList([ResnetBlock(a, mid), Upsample(mid, b)]))
        self._hyper_rows(None)

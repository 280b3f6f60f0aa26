"""Context networks of the eps variant (reference epsilonparam/modules/compress_modules.py)."""
from torch import nn

from cdc_compression_b200._shared.compressor_impl import HyperpriorCompressor
from .network_components import GDN1, Downsample, ResnetBlock, Upsample, VBRCondition


class Compressor(HyperpriorCompressor):
    has_vbr_slot = True

    def __init__(self, dim=64, dim_mults=(1, 2, 3, 3), hyper_dims_mults=(3, 3, 3), channels=3, out_channels=3,
                 vbr=False):
        super().__init__()
        self._init_dims(dim, dim_mults, list(reversed([out_channels] + [dim * m for m in dim_mults])),
                        hyper_dims_mults, channels, out_channels)
        self.vbr = vbr

    def _vbr(self, width):
        return VBRCondition(1, width) if self.vbr else nn.Identity()


class BigCompressor(Compressor):
    """ResnetBlock + strided-conv analysis, ResnetBlock + transposed-conv synthesis (the demo's context_fn)."""

    def __init__(self, dim=64, dim_mults=(1, 3, 3, 3), hyper_dims_mults=(3, 3, 3), channels=3, out_channels=3,
                 vbr=False):
        super().__init__(dim, dim_mults, hyper_dims_mults, channels, out_channels, vbr)
        self.build_network()

    def build_network(self):
        super().build_network()
        for i, (a, b) in enumerate(self.in_out):
            self.enc.append(nn.ModuleList([ResnetBlock(a, b, None, i == 0), self._vbr(b), Downsample(b)]))
        n = len(self.reversed_in_out)
        for i, (a, b) in enumerate(self.reversed_in_out):
            mid = a if i >= n - 1 else b          # the last stage keeps its width until the upsampler
            self.dec.append(nn.ModuleList([ResnetBlock(a, mid), self._vbr(mid), Upsample(mid, b)]))
        self._hyper_rows(self._vbr)


class SimpleCompressor(Compressor):
    """conv5/GDN analysis and synthesis (imported by the demo script, not used by it)."""

    def __init__(self, dim=64, dim_mults=(1, 2, 3, 3), hyper_dims_mults=(3, 3, 3), channels=3, out_channels=3,
                 vbr=False):
        super().__init__(dim, dim_mults, hyper_dims_mults, channels, out_channels, vbr)
        self.build_network()

    def build_network(self):
        super().build_network()
        n = len(self.in_out)
        for i, (a, b) in enumerate(self.in_out):
            last = i >= n - 1
            self.enc.append(nn.ModuleList([nn.Conv2d(a, b, 5, 2, 2), self._vbr(b) if not last else nn.Identity(),
                                           nn.Identity() if last else GDN1(b)]))
        n = len(self.reversed_in_out)
        for i, (a, b) in enumerate(self.reversed_in_out):
            last = i >= n - 1
            self.dec.append(nn.ModuleList([nn.ConvTranspose2d(a, b, 5, 2, 2, 1),
                                           self._vbr(b) if not last else nn.Identity(),
                                           nn.Identity() if last else GDN1(b, True)]))
        self._hyper_rows(self._vbr)

"""cdc_compression_b200 — B200-native denoiser for conditional-diffusion image decompression.

The package holds exactly what the decoder hot path needs:

* ``csrc/``            hand-written sm_100a CUDA kernels + the C-ABI engine (``libcdc_b200.so``)
* ``_native.py``       ctypes binding of ``include/cdc_b200.h``
* ``engine.py``        thin Python owner of a ``cdc_engine`` (tensor ownership stays with PyTorch)
* ``epsilonparam/``, ``xparam/``  drop-in mirrors of the reference's ``modules`` packages
  (same class names, constructor kwargs and ``state_dict`` keys), whose ``Unet.forward`` and
  DDIM loop run on the engine.  There is no CPU / eager fallback on the product path.
"""
from .engine import DenoiserEngine, EngineError, native_available  # noqa: F401

__all__ = ["DenoiserEngine", "EngineError", "native_available"]

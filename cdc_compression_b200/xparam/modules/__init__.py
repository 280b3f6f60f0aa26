"""Drop-in ``modules`` package: same import surface as the reference's, backed by the CUDA engine.

The reference's demo scripts run with this variant directory as cwd (``from modules.unet import Unet``);
make the repository root importable so the shared implementation (``cdc_compression_b200``) resolves.
"""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

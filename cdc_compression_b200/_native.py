"""ctypes binding of include/cdc_b200.h (libcdc_b200.so).  Fails loudly if the library is absent."""
from __future__ import annotations

import ctypes as C
import os

_LIB_NAME = "libcdc_b200.so"
_lib = None

CDC_ABI_VERSION = 1
CDC_MAX_LEVELS = 8
VARIANT = {"eps": 0, "x": 1}
CLIP = {"none": 0, "full": 1, "half": 2}
PRED = {"noise": 0, "x": 1, "v": 2}


class CdcConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("variant", C.c_int32), ("dim", C.c_int32), ("channels", C.c_int32),
        ("context_channels", C.c_int32), ("n_levels", C.c_int32), ("dim_mults", C.c_int32 * CDC_MAX_LEVELS),
        ("n_context", C.c_int32), ("context_dim_mults", C.c_int32 * CDC_MAX_LEVELS),
    ]


class CdcStepCoef(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "sqrt_recip_acp", "sqrt_recipm1_acp", "sqrt_acp_prev", "dir_coef", "noise_coef", "unet_time",
        "sqrt_acp", "sqrt_1m_acp")]


# name -> (restype, argtypes); mirrors include/cdc_b200.h one to one
_P = C.c_void_p
_SIGNATURES = {
    "cdc_abi_version": (C.c_int, []),
    "cdc_last_error": (C.c_char_p, [_P]),
    "cdc_engine_create": (C.c_int, [C.POINTER(CdcConfig), C.c_int, C.POINTER(_P)]),
    "cdc_engine_destroy": (None, [_P]),
    "cdc_engine_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "cdc_engine_finalize": (C.c_int, [_P]),
    "cdc_engine_workspace_bytes": (C.c_int64, [_P, C.c_int, C.c_int, C.c_int]),
    "cdc_unet_forward": (C.c_int, [_P, _P, _P, C.POINTER(_P), C.c_int, _P, C.c_int, C.c_int, C.c_int, _P,
                                   C.c_int64, _P]),
    "cdc_set_context": (C.c_int, [_P, C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int64, _P]),
    "cdc_engine_has_context_decoder": (C.c_int, [_P]),
    "cdc_context_decode": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int64, _P]),
    "cdc_engine_read_context": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int64, _P]),
    "cdc_set_schedule": (C.c_int, [_P, C.POINTER(CdcStepCoef), C.c_int, _P]),
    "cdc_ddim_step": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P,
                                C.c_int64, _P]),
    "cdc_sample_loop": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P,
                                  C.c_int64, _P]),
    "cdc_sample_loop_noise": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P,
                                        C.c_int64, _P]),
    "cdc_engine_launches_per_forward": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "cdc_engine_launches_per_step": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "cdc_engine_flops_per_forward": (C.c_double, [_P, C.c_int, C.c_int, C.c_int]),
    "cdc_engine_debug_read": (C.c_int64, [_P, C.c_int, _P, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int)]),
    "cdc_engine_num_ops": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "cdc_engine_op_name": (C.c_char_p, [_P, C.c_int]),
    "cdc_engine_set_debug": (C.c_int, [_P, C.c_int]),
    "cdc_engine_profile_ops": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P]),
    "cdc_engine_set_mainloop": (C.c_int, [_P, C.c_int]),
    "cdc_engine_tc_ops": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
}


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """dlopen the engine; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). cdc_compression_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.cdc_abi_version() != CDC_ABI_VERSION:
        raise ImportError("libcdc_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib

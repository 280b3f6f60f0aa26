"""Helpers with the reference's names (epsilonparam/modules/utils.py)."""
from inspect import isfunction

from cdc_compression_b200._shared.diffusion_impl import cosine_beta_schedule, extract, linear_beta_schedule  # noqa: F401
from cdc_compression_b200._shared.layers import dequantize, normal_box_likelihood  # noqa: F401


def exists(x):
    return x is not None


def default(val, d):
    if val is not None:
        return val
    return d() if isfunction(d) else d

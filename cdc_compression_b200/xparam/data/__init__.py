"""Placeholder for the reference's training-data package (`from data import load_data` in the demo
script is an unused import).  Training data loading is outside the decoder hot path."""


def load_data(*args, **kwargs):
    raise NotImplementedError("cdc_compression_b200 ships the decoder hot path only; use the reference's "
                              "`data` package for training datasets")

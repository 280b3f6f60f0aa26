"""Minimal stand-in for ``ema_pytorch.EMA`` (0.2.3), used by the x-variant demo script only as a
checkpoint container: ``ema = EMA(diffusion, ...); ema.load_state_dict(ckpt["ema"]); ema.ema_model``.

state_dict layout: ``online_model.*``, ``ema_model.*`` and the buffers ``initted`` / ``step``.  The
``ema_model.`` prefix is evidenced in the reference tree (epsilonparam/modules/distill_trainer.py:101-107);
the remaining names follow the upstream package and are to be re-checked against a real x-param
checkpoint when one is available (none is reachable offline — SURVEY.md §8b).  If the real package
is installed it is imported instead of this file only when it precedes the script directory on sys.path.
"""
import copy

import torch
from torch import nn


class EMA(nn.Module):
    def __init__(self, model, ema_model=None, beta=0.9999, update_after_step=100, update_every=10, inv_gamma=1.0,
                 power=2 / 3, min_value=0.0, param_or_buffer_names_no_ema=(), ignore_names=(),
                 ignore_startswith_names=()):
        super().__init__()
        self.beta = beta
        self.online_model = model
        self.ema_model = ema_model if ema_model is not None else copy.deepcopy(model)
        self.ema_model.requires_grad_(False)
        self.update_every = update_every
        self.update_after_step = update_after_step
        self.inv_gamma = inv_gamma
        self.power = power
        self.min_value = min_value
        self.register_buffer("initted", torch.Tensor([False]))
        self.register_buffer("step", torch.tensor([0]))

    def update(self):
        raise NotImplementedError("EMA updates belong to training, which is outside the decoder hot path")

    def forward(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)
